#!/usr/bin/env python
"""Per-stage CUDA-event timings of one walker step of any BASELINE workload (development aid; bench.py is the contract).
  python tools/stage_times.py [--workload c_ecp] [--walkers 65536] [--reps 5] [--lib path/to/alt.so] [--stages sweep,base,quad,final]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c_ecp")
ap.add_argument("--walkers", type=int, default=0)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--lib", default=None)
ap.add_argument("--stages", default="sweep,base,quad,final")
args = ap.parse_args()
if args.lib:
    os.environ["AIQMC_LIB"] = args.lib
import numpy as np
import torch
import aiqmc_b200
from aiqmc_b200 import workloads as W

wl = W.build(args.workload, args.walkers or None)
B = wl.pos.shape[0]
eng = wl.engine()
g1, g2c, u = eng.rng_sweep(1, 0, 0, B, W.TSTEP)
rot = eng.rng_rotations(1, 0, 0, B) if wl.ecp is not None else None
pos = torch.from_numpy(wl.pos.copy()).cuda()
e_l = torch.empty((B, 2), dtype=torch.float64, device="cuda")
seed_a = torch.randn(B, dtype=torch.float64, device="cuda")
seed_b = torch.randn(B, dtype=torch.float64, device="cuda")
if wl.ecp is not None:
    ops = {"sweep": lambda: eng.vmc_sweep(pos, g1, g2c, u, W.TSTEP, want_accept=False),
           "base": lambda: eng.local_energy(pos, rot, stages=1, out=e_l),
           "quad": lambda: eng.local_energy(pos, rot, stages=2, out=e_l),
           "quad_coop": lambda: eng.local_energy(pos, rot, stages=2 | 16, out=e_l),
           "final": lambda: eng.local_energy(pos, rot, stages=4, out=e_l),
           "grad": lambda: eng.psi(pos, mode=1),
           "lap": lambda: eng.psi(pos, mode=2),
           "pgrad": lambda: eng.param_grad(pos, seed_a, seed_b)}
    ops["base"]()          # fills the workspace the quadrature stage reads (v_l tables, move cache)
else:
    ops = {"sweep": lambda: eng.vmc_sweep(pos, g1, g2c, u, W.TSTEP, want_accept=False),
           "base": lambda: eng.local_energy(pos), "grad": lambda: eng.psi(pos, mode=1), "lap": lambda: eng.psi(pos, mode=2)}
    args.stages = ",".join(s for s in args.stages.split(",") if s in ops)
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
print(f"step       {tot:9.3f} ms  -> {B / tot * 1e3:,.0f} walker-steps/s  ({args.workload}, {B} walkers)")
