"""Host-side multi-GPU logic (aiqmc_b200.parallel) with world_size 2 over gloo on CPU: walker sharding, the
4-double energy all-reduce, the MIN all-reduce of quirk Q20 and the cross-rank population control.  On the GPU
box the same functions run over NCCL with the CUDA comb / gather kernels plugged in."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from common import O

import aiqmc_b200
from aiqmc_b200 import parallel


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, B, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(7)                       # same stream on every rank: the "global" arrays
        e_all = torch.tensor(rng.normal(-5.4, 0.3, size=world * B) + 1j * rng.normal(0, 0.01, size=world * B))
        w_all = torch.tensor(rng.uniform(0.2, 1.8, size=world * B))
        p_all = torch.tensor(rng.normal(size=(world * B, 12)))
        lo, hi = parallel.shard_bounds(world * B, rank, world)
        assert (lo, hi) == (rank * B, (rank + 1) * B)
        e = e_all[lo:hi]
        # energy statistics: local partials -> one 4-double all-reduce
        stats = torch.stack([e.real.sum(), e.imag.sum(), (e.abs() ** 2).sum(), torch.tensor(float(B), dtype=torch.float64)])
        mean, var, cnt = parallel.allreduce_energy_stats(stats)
        assert int(cnt) == world * B
        np.testing.assert_allclose(complex(mean), complex(e_all.mean()), rtol=1e-13)
        np.testing.assert_allclose(float(var), float((e_all.abs() ** 2).mean() - abs(complex(e_all.mean())) ** 2), rtol=1e-10)
        # Q20 global minimum
        m = parallel.allreduce_min(e.real.min().reshape(1).clone())
        assert float(m) == float(e_all.real.min())
        # global comb + migration against the single-process oracle on the concatenated arrays
        neww, newp, src, imported, moved = parallel.global_branch(w_all[lo:hi], p_all[lo:hi].contiguous(), 0.37)
        neww_ref, inds_ref = O.branch(w_all, 0.37)
        np.testing.assert_allclose(float(neww), float(neww_ref), rtol=1e-15)
        assert np.array_equal(src.numpy(), inds_ref.numpy()[lo:hi] // B)          # source rank of every new walker
        np.testing.assert_array_equal(newp.numpy(), p_all.numpy()[inds_ref.numpy()[lo:hi]])
        assert imported == int(((inds_ref[lo:hi] < lo) | (inds_ref[lo:hi] >= hi)).sum())
        # only the migrating walkers were sent: what this rank exported = what the other rank imported from it
        exported = int(((inds_ref // B == rank) & (torch.arange(world * B) // B != rank)).sum())
        assert moved == exported * 12 * 8 and moved < B * 12 * 8 * world
        # loss side (pploss.py:73-135 across ranks): pmean'ed clipping statistics, all-gathered median, pmean(grad)
        g_local = torch.tensor(rng.normal(size=5)) * (rank + 1)
        np.testing.assert_allclose(parallel.allreduce_mean(g_local).numpy(),
                                   sum(g_local.numpy() / (rank + 1) * (r + 1) for r in range(world)) / world, rtol=1e-14)
        assert torch.equal(parallel.all_gather_cat(e.real.contiguous()), e_all.real)
        for median in (True, False):
            centre, diff = aiqmc_b200.clip_local_values(e, e_all.mean(), 1.5, median, True, complex_output=True)
            centre_ref, diff_ref = O.clip_local_values(e_all, e_all.mean(), 1.5, median, True, complex_output=True)
            np.testing.assert_allclose(complex(centre), complex(centre_ref), rtol=1e-12)
            np.testing.assert_allclose(diff.numpy(), diff_ref.numpy()[lo:hi], rtol=1e-12, atol=1e-14)
        out[rank] = 1
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_world_size_2_gloo():
    world, B = 2, 257
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, B, out), nprocs=world, join=True)
    assert dict(out) == {0: 1, 1: 1}


def test_shard_bounds_cover_everything_once():
    for n, world in [(10, 3), (65536, 8), (7, 8), (0, 2)]:
        cuts = [parallel.shard_bounds(n, r, world) for r in range(world)]
        assert cuts[0][0] == 0 and cuts[-1][1] == n
        assert all(cuts[r][1] == cuts[r + 1][0] for r in range(world - 1))
        sizes = [hi - lo for lo, hi in cuts]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        parallel.shard_bounds(10, 3, 3)


def test_single_process_global_branch_is_the_local_comb():
    rng = np.random.default_rng(3)
    w = torch.tensor(rng.uniform(0.1, 2.0, size=100))
    p = torch.tensor(rng.normal(size=(100, 6)))
    neww, newp, src, imported, moved = parallel.global_branch(w, p, 0.9)
    neww_ref, inds_ref = O.branch(w, 0.9)
    assert imported == 0 and moved == 0 and abs(float(neww) - float(neww_ref)) < 1e-15
    np.testing.assert_array_equal(newp.numpy(), p.numpy()[inds_ref.numpy()])
