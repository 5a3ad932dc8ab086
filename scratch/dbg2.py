import sys, os; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch
from common import CASES, Case, O
import aiqmc_b200
print("lib", os.environ.get("AIQMC_LIB"))
case = Case(**CASES["N2_ecp"], nwalkers=4)
eng = aiqmc_b200.WalkerEngine(case.spec(), case.params)
f = lambda x: case.net.apply(case.params, x, case.t_spins, case.t_atoms)[1]
_, gt, dt = O.grad_and_hess_diag(f, torch.tensor(case.pos))
for mode in (1,2):
    out = eng.psi(case.pos, mode=mode)
    g = out[2].cpu().numpy()
    bad = np.abs(g-gt.numpy())>1e-6
    print("mode",mode,"bad idx row0:", np.nonzero(bad[0])[0])
