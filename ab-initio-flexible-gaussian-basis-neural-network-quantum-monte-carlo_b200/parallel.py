"""Multi-GPU plumbing of the walker path: one process per GPU, walkers sharded contiguously, parameters
replicated.  The data path has exactly three exchanges:

  * energy mean / variance: ONE 4-double all-reduce of [sum Re E, sum Im E, sum |E|^2, count]
    (the two pmean's of Loss/pploss.py:165-167, DMC/total_energy.py:28-30);
  * the DMC e_cut of quirk Q20: a 1-double MIN all-reduce (DMC/S_matrix.py:4-25);
  * DMC population control across GPUs (new capability; the reference combs per device,
    DMC/branch.py:10-34 inside pmap): ONE weight total per rank is all-gathered, every rank scans those totals the
    same way, locates every tooth of the GLOBAL systematic comb rank-first and its own teeth walker-exact, and a
    single uneven all-to-all moves only the walkers that change rank.  On the GPU this schedule runs
    behind the C ABI (csrc/population.cu: aiqmc_rebalance_nccl over raw ncclSend/ncclRecv, block totals of a blocked
    scan instead of rank totals so that the result is bit-identical to the single-GPU comb); `global_branch` below
    is its device-agnostic torch statement, driven with world_size 2 over gloo by tests/test_distributed_gloo.py.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of n_total walkers for `rank`; sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _world(group) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def allreduce_energy_stats(stats: torch.Tensor, group=None):
    """stats = [sum Re E, sum Im E, sum |E|^2, count] (device) -> (mean complex, variance, count), global."""
    _, world = _world(group)
    if world > 1:
        dist.all_reduce(stats, group=group)
    cnt = stats[3]
    mean = torch.complex(stats[0], stats[1]) / cnt
    variance = stats[2] / cnt - (mean.real ** 2 + mean.imag ** 2)
    return mean, variance, cnt


def allreduce_mean(x: torch.Tensor, group=None) -> torch.Tensor:
    """constants.pmean (kfac_jax pmean_if_pmap): mean over the ranks of `group`; identity for a single process."""
    _, world = _world(group)
    if world > 1:
        x = x.clone()
        dist.all_reduce(x, group=group)
        x = x / world
    return x


def all_gather_cat(x: torch.Tensor, group=None) -> torch.Tensor:
    """constants.all_gather flattened: every rank's (B,) shard concatenated (clip_from_median, pploss.py:118)."""
    _, world = _world(group)
    if world == 1:
        return x
    parts = [torch.empty_like(x) for _ in range(world)]
    dist.all_gather(parts, x.contiguous(), group=group)
    return torch.cat(parts)


def allreduce_min(x: torch.Tensor, group=None) -> torch.Tensor:
    _, world = _world(group)
    if world > 1:
        dist.all_reduce(x, op=dist.ReduceOp.MIN, group=group)
    return x


def global_branch(weights: torch.Tensor, positions: torch.Tensor, u: float, group=None):
    """Cross-rank systematic comb + walker migration, device-agnostic torch statement of the schedule the CUDA/NCCL
    path runs (csrc/population.cu: aiqmc_rebalance_nccl).  Exchanged: ONE weight total per rank
    (all-gather), then only the walkers that change rank (all_to_all_single with uneven splits); the weights and the
    positions of the other ranks are never gathered.

    weights (B,), positions (B,3N): this rank's shard (equal B on every rank); u in [0,1): the comb's uniform
    (DMC/branch.py:21).  Slot k of rank r receives the walker that tooth r*B + k of the GLOBAL comb selects.
    Returns (new_weight, new_positions (B,3N), source rank of every new walker (B,), walkers imported from other
    ranks, bytes sent to other ranks).  tests/test_distributed_gloo.py drives it with world_size 2 over gloo."""
    rank, world = _world(group)
    B = weights.shape[0]
    Bt = world * B
    dev = weights.device
    cum = torch.cumsum(weights, 0)
    totals = cum[-1:].clone()
    if world > 1:
        parts = [torch.empty_like(totals) for _ in range(world)]
        dist.all_gather(parts, totals, group=group)
        totals = torch.cat(parts)
    off = torch.zeros(world + 1, dtype=weights.dtype, device=dev)
    for r in range(world):                                    # sequential: the same association on every rank
        off[r + 1] = off[r] + totals[r]
    wtot = off[world]
    k = torch.arange(Bt, device=dev)
    v = torch.remainder(u * wtot + k.to(weights.dtype) * (wtot / Bt), wtot)
    src_rank = torch.searchsorted(off[1:].contiguous(), v, right=False).clamp_(max=world - 1)
    mine = src_rank == rank
    local = torch.searchsorted((off[rank] + cum).contiguous(), v[mine], right=False).clamp_(max=B - 1)
    dest = torch.div(k[mine], B, rounding_mode="floor")
    rows = positions[local].contiguous()                      # tooth order = grouped by destination, ascending
    send_counts = torch.bincount(dest, minlength=world).tolist()
    my_src = src_rank[rank * B:(rank + 1) * B]
    recv_counts = torch.bincount(my_src, minlength=world).tolist()
    if world > 1:
        got = torch.empty((B,) + tuple(positions.shape[1:]), dtype=positions.dtype, device=dev)
        dist.all_to_all_single(got, rows, output_split_sizes=recv_counts, input_split_sizes=send_counts, group=group)
    else:
        got = rows
    order = torch.sort(my_src, stable=True).indices           # messages arrive grouped by source, each in tooth order
    new_pos = torch.empty_like(got)
    new_pos[order] = got
    row_bytes = positions[0].numel() * positions.element_size() if B else 0
    moved = (sum(send_counts) - send_counts[rank]) * row_bytes
    return wtot / Bt, new_pos, my_src, int((my_src != rank).sum()), int(moved)
