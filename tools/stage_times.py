#!/usr/bin/env python
"""Per-launcher CUDA-event timings of one walker step (development aid; bench.py is the contract).
  python tools/stage_times.py [--walkers 65536] [--reps 5] [--lib path/to/alt.so]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
ap = argparse.ArgumentParser()
ap.add_argument("--walkers", type=int, default=65536)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--lib", default=None)
ap.add_argument("--stages", default="sweep,base,quad,final")
args = ap.parse_args()
if args.lib:
    os.environ["AIQMC_LIB"] = args.lib
import numpy as np
import torch
import bench
import aiqmc_b200

B = args.walkers
case, tabs = bench.build_case(B)
eng = aiqmc_b200.WalkerEngine(case.spec(), case.params, ecp=aiqmc_b200.make_ecp(1, list_l=2, **tabs))
rng = np.random.default_rng(5)
r = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in bench.make_rand(rng, B, case.n, bench.TSTEP).items()}
rot = torch.from_numpy(bench.random_rot(rng, B)).cuda()
pos = torch.from_numpy(case.pos.copy()).cuda()
e_l = torch.empty((B, 2), dtype=torch.float64, device="cuda")
ops = {
    "sweep": lambda: eng.vmc_sweep(pos, r["gauss1"], r["gauss2"], r["rnd"], bench.TSTEP, want_accept=False),
    "base": lambda: eng.local_energy(pos, rot, stages=1, out=e_l),
    "quad": lambda: eng.local_energy(pos, rot, stages=2, out=e_l),
    "quad_coop": lambda: eng.local_energy(pos, rot, stages=2 | 16, out=e_l),
    "final": lambda: eng.local_energy(pos, rot, stages=4, out=e_l),
    "pgrad": lambda: eng.param_grad(pos, seed_a, seed_b),          # loss-gradient side, not part of a walker step
}
seed_a = torch.randn(B, dtype=torch.float64, device="cuda")
seed_b = torch.randn(B, dtype=torch.float64, device="cuda")
ops["base"]()          # fills the workspace the quadrature stage reads (v_l tables, move cache)
tot = 0.0
for name in args.stages.split(","):
    f = ops[name]
    for _ in range(2):
        f()
    torch.cuda.synchronize()
    ts = []
    for _ in range(args.reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = float(np.median(ts))
    tot += ms if name in ("sweep", "base", "quad", "final") else 0.0
    print(f"{name:10s} {ms:9.3f} ms")
print(f"step       {tot:9.3f} ms  -> {B / tot * 1e3:,.0f} walker-steps/s")
