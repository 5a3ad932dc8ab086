#!/usr/bin/env python
"""Run under torchrun on N GPUs of one box: checks the multi-GPU exchanges of the walker path against single-GPU
recomputation -- the 4-double energy all-reduce and the Q20 MIN all-reduce (through torch.distributed AND through the
C ABI: aiqmc_energy_allreduce / aiqmc_ecut_allreduce_min), and the cross-GPU population control (aiqmc_rebalance_nccl:
block totals all-gathered, only the migrating walkers sent) against the single-GPU comb + gather over the concatenated
batch (bit for bit) and against the torch statement of the same schedule (parallel.global_branch over NCCL).
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multigpu_check.py"""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import aiqmc_b200  # noqa: E402
from aiqmc_b200 import parallel, workloads as W  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
lines = []
for name, B in (("c_ecp", 65536), ("c6h6", 4096)):
    wl = W.build(name, 2)
    eng = wl.engine(dev)
    row = 3 * wl.n
    rng = np.random.default_rng(11)                         # the same "global" arrays on every rank
    w_all = torch.tensor(rng.uniform(0.1, 2.0, size=world * B))
    p_all = torch.tensor(rng.normal(size=(world * B, row)))
    e_all = torch.tensor(rng.normal(-5.4, 0.3, size=world * B) + 1j * rng.normal(0, 0.01, size=world * B))
    lo, hi = parallel.shard_bounds(world * B, rank, world)
    w, p, e = w_all[lo:hi].to(dev), p_all[lo:hi].to(dev).contiguous(), e_all[lo:hi].to(dev)

    comm = aiqmc_b200.NcclComm.for_group(None, dev)
    # energy statistics: torch.distributed and the C-ABI collective agree with the global mean
    stats = eng.energy_stats(e)
    stats_c = stats.clone()
    mean, var, cnt = parallel.allreduce_energy_stats(stats)
    aiqmc_b200.lib.check(eng.lib.aiqmc_energy_allreduce(C.c_void_p(stats_c.data_ptr()), comm.handle,
                                                        C.c_void_p(torch.cuda.current_stream().cuda_stream)), "aiqmc_energy_allreduce")
    assert int(cnt) == world * B and torch.allclose(stats, stats_c, rtol=1e-13, atol=0)
    np.testing.assert_allclose(complex(mean), complex(e_all.mean()), rtol=1e-12)
    bc = torch.full((B,), 3.0, dtype=torch.float64, device=dev)
    m = eng.dmc_ecut_min(e, -5.41, bc)
    m_c = m.clone()
    parallel.allreduce_min(m)
    aiqmc_b200.lib.check(eng.lib.aiqmc_ecut_allreduce_min(C.c_void_p(m_c.data_ptr()), comm.handle,
                                                          C.c_void_p(torch.cuda.current_stream().cuda_stream)), "aiqmc_ecut_allreduce_min")
    want = min(float((-5.41 - e_all.real).abs().min()), 3.0)
    np.testing.assert_allclose(float(m), want, rtol=1e-13)
    assert float(m_c) == float(m)

    # population control through the C ABI vs one GPU doing the whole batch
    neww, newp, src, imported, moved = aiqmc_b200.branch_global(eng, w, p, 0.37, return_bytes=True, mode="ordered")
    neww1, inds1 = eng.branch_comb(w_all.to(dev), 0.37)
    ref = p_all.to(dev)[inds1[lo:hi].long()]
    assert float(neww) == float(neww1), (float(neww), float(neww1))
    assert torch.equal(newp, ref), "distributed comb differs from the single-GPU comb"
    assert torch.equal(src.long(), inds1[lo:hi].long() // B)
    # ... and vs the torch statement of the same schedule over NCCL (rank totals instead of block totals)
    neww2, newp2, src2, imported2, moved2 = parallel.global_branch(w, p, 0.37)
    same = float((newp2 == newp).all(dim=1).double().mean())
    tot = torch.tensor([imported, moved], device=dev, dtype=torch.float64)
    dist.all_reduce(tot)
    # balanced mode: the same MULTISET of walkers as the single-GPU comb, but only the population imbalance moves
    nb_w, nb_p, nb_src, nb_imp, nb_moved = aiqmc_b200.branch_global(eng, w, p, 0.37, return_bytes=True, mode="balanced")
    assert float(nb_w) == float(neww1)
    allp = [torch.empty_like(nb_p) for _ in range(world)]
    dist.all_gather(allp, nb_p.contiguous())
    got_rows = torch.cat(allp).cpu().numpy()
    ref_rows = p_all.numpy()[inds1.cpu().numpy().astype(np.int64)]
    key_cols = tuple(got_rows[:, c] for c in range(min(3, row) - 1, -1, -1))
    ref_cols = tuple(ref_rows[:, c] for c in range(min(3, row) - 1, -1, -1))
    assert np.array_equal(got_rows[np.lexsort(key_cols)], ref_rows[np.lexsort(ref_cols)]), "balanced mode changed the multiset"
    tb = torch.tensor([nb_imp, nb_moved], device=dev, dtype=torch.float64)
    dist.all_reduce(tb)
    gathered_everything = world * B * row * 8 * (world - 1)          # what round 1 moved: every rank received all positions
    if rank == 0:
        lines.append(f"{name}: world {world}, {world * B} walkers x {row * 8} B: {int(tot[0])} walkers migrated, "
                     f"{tot[1] / 1e6:.2f} MB sent in total ({tot[1] / world / 1e6:.2f} MB per rank; own shard {B * row * 8 / 1e6:.2f} MB; "
                     f"the all-gather of round 1 moved {gathered_everything / 1e6:.1f} MB), new weight {float(neww):.6f}, "
                     f"bit-identical to the single-GPU comb; torch/rank-totals schedule agrees on {same * 100:.4f} % of the rows; "
                     f"BALANCED mode: same multiset of walkers, {int(tb[0])} walkers migrated, {tb[1] / 1e6:.3f} MB sent in total")
if rank == 0:
    print("multigpu_check ok")
    for ln in lines:
        print(ln)
dist.destroy_process_group()
