#!/usr/bin/env python
"""Run under torchrun on N GPUs of one box: checks the NCCL paths of aiqmc_b200.parallel against single-GPU
recomputation -- energy all-reduce, Q20 MIN all-reduce, cross-GPU population control (global comb + migration).
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multigpu_check.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import aiqmc_b200  # noqa: E402
from aiqmc_b200 import parallel  # noqa: E402
from common import CASES, Case  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B = 4096
case = Case(**CASES["C_ecp"], nwalkers=2)
eng = aiqmc_b200.WalkerEngine(case.spec(), case.params, device=dev)
rng = np.random.default_rng(11)                         # the same "global" arrays on every rank
w_all = torch.tensor(rng.uniform(0.1, 2.0, size=world * B))
p_all = torch.tensor(rng.normal(size=(world * B, 12)))
e_all = torch.tensor(rng.normal(-5.4, 0.3, size=world * B) + 1j * rng.normal(0, 0.01, size=world * B))
lo, hi = parallel.shard_bounds(world * B, rank, world)
w, p, e = w_all[lo:hi].to(dev), p_all[lo:hi].to(dev).contiguous(), e_all[lo:hi].to(dev)

mean, var, cnt = parallel.allreduce_energy_stats(eng.energy_stats(e))
assert int(cnt) == world * B
np.testing.assert_allclose(complex(mean), complex(e_all.mean()), rtol=1e-12)
m = parallel.allreduce_min(eng.dmc_ecut_min(e, -5.41, torch.full((B,), 3.0, dtype=torch.float64, device=dev)))
want = min(float((-5.41 - e_all.real).abs().min()), 3.0)
np.testing.assert_allclose(float(m), want, rtol=1e-13)

neww, newp, src, imported = aiqmc_b200.branch_global(eng, w, p, 0.37)
# single-GPU recomputation of the global comb with the same CUDA kernels
neww1, inds1 = eng.branch_comb(w_all.to(dev), 0.37)
assert float(neww) == float(neww1)
assert torch.equal(src, inds1[lo:hi])
assert torch.equal(newp, p_all.to(dev)[inds1[lo:hi].long()])
tot = torch.tensor([imported], device=dev)
dist.all_reduce(tot)
if rank == 0:
    print(f"multigpu_check ok: world {world}, {world * B} walkers, {int(tot)} migrated across ranks, new weight {float(neww):.6f}")
dist.destroy_process_group()
