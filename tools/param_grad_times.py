import sys, os, time, math
import numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import aiqmc_b200
from common import Case
ring = lambda r, n: [[r * math.cos(2 * math.pi * k / n), r * math.sin(2 * math.pi * k / n), 0.0] for k in range(n)]
for name, kw, B in (("C", dict(n=4, natoms=1, spins=[1., -1., 1., -1.], atoms=[[0., 0., 0.]], charges=[4.0]), 65536),
                    ("N2", dict(n=10, natoms=2, spins=[1.] * 5 + [-1.] * 5, atoms=[[0, 0, -1.034], [0, 0, 1.034]], charges=[5.0, 5.0]), 65536),
                    ("C6H6", dict(n=30, natoms=12, spins=[1.] * 15 + [-1.] * 15, atoms=ring(2.640, 6) + ring(4.689, 6), charges=[4.0] * 6 + [1.0] * 6), 2368)):
    case = Case(seed=1, nwalkers=B, width=1.0, **kw)
    eng = aiqmc_b200.WalkerEngine(case.spec(), case.params)
    pos = torch.tensor(case.pos).cuda()
    a = torch.randn(B, dtype=torch.float64, device="cuda"); b = torch.randn(B, dtype=torch.float64, device="cuda")
    eng.param_grad(pos, a, b); torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.param_grad(pos, a, b); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    print(f"param_grad {name:5s} B={B:6d} P={eng.layout.total:6d}  {np.median(ts):9.3f} ms", flush=True)
