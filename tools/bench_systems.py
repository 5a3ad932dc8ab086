#!/usr/bin/env python
"""Walker-step timings (sweep + local energy, CUDA events) for the other BASELINE.json systems; development aid --
bench.py measures configs[1] only.   python tools/bench_systems.py [N2 C6H6 C_ae]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import aiqmc_b200  # noqa: E402
import bench  # noqa: E402
from common import Case, ecp_tables  # noqa: E402

ring = lambda r, n, ph=0.0: [[r * np.cos(2 * np.pi * k / n + ph), r * np.sin(2 * np.pi * k / n + ph), 0.0] for k in range(n)]
SYSTEMS = {
    "C_ecp": dict(n=4, natoms=1, spins=[1., -1., 1., -1.], atoms=[[0., 0., 0.]], charges=[4.0], B=65536, ecp=True),
    "C_ae": dict(n=6, natoms=1, spins=[1.] * 3 + [-1.] * 3, atoms=[[0., 0., 0.]], charges=[6.0], B=4096, ecp=False),
    "N2": dict(n=10, natoms=2, spins=[1.] * 5 + [-1.] * 5, atoms=[[0, 0, -1.034], [0, 0, 1.034]], charges=[5.0, 5.0],
               B=16384, ecp=True),
    "C6H6": dict(n=30, natoms=12, spins=[1.] * 15 + [-1.] * 15, atoms=ring(2.640, 6) + ring(4.689, 6),
                 charges=[4.0] * 6 + [1.0] * 6, B=256, ecp=True),
}
for name in (sys.argv[1:] or ["N2", "C_ae", "C6H6"]):
    s = dict(SYSTEMS[name])
    B, with_ecp = s.pop("B"), s.pop("ecp")
    B = int(os.environ.get("WALKERS", B))
    case = Case(seed=20260101, nwalkers=B, width=1.0, **s)
    case.params = case.net.init(np.random.default_rng(1), randomize_all=False)
    ecp = aiqmc_b200.make_ecp(case.a, list_l=2, **ecp_tables(case.a)) if with_ecp else None
    eng = aiqmc_b200.WalkerEngine(case.spec(), case.params, ecp=ecp)
    rng = np.random.default_rng(5)
    r = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in bench.make_rand(rng, B, case.n, bench.TSTEP).items()}
    rot = torch.from_numpy(bench.random_rot(rng, B)).cuda() if with_ecp else None
    pos = torch.from_numpy(case.pos.copy()).cuda()
    out = {}
    stages = [("sweep", lambda: eng.vmc_sweep(pos, r["gauss1"], r["gauss2"], r["rnd"], bench.TSTEP, want_accept=False)),
              ("energy", lambda: eng.local_energy(pos, rot))]
    if with_ecp and os.environ.get("STAGES"):       # energy split: 1 = kinetic + Coulomb + local channel, 2 = quadrature
        stages += [("e:base", lambda: eng.local_energy(pos, rot, stages=1)), ("e:quad", lambda: eng.local_energy(pos, rot, stages=2))]
    for label, f in stages:
        t0 = time.time()
        f(); torch.cuda.synchronize()
        first = time.time() - t0
        ts = []
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); f(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        out[label] = float(np.median(ts))
        print(f"{name:6s} B={B:6d} {label:7s} {out[label]:10.3f} ms   (first call {first:.2f} s)", flush=True)
    tot = out["sweep"] + out["energy"]
    flops = (bench.flops_walker_step_ecp(case.n, case.a) if with_ecp else (6 * case.n + 5) * bench.flops_psi(case.n, case.a))
    print(f"{name:6s} step {tot:10.3f} ms -> {B / tot * 1e3:12,.0f} walker-steps/s, {B / tot * 1e3 * flops / 1e12:6.2f} algorithmic TFLOP/s", flush=True)
