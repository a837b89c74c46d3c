"""Multi-GPU sharding of the read index space and the path's only collective.

Reads are independent (bernoullimodule.c:182-263 keeps no state across reads), so rank g of G gets
the contiguous chunk [g*N/G, (g+1)*N/G): concatenating the per-rank outputs in rank order restores
the reference's sequential order (moira.py:455-487).  After the last batch the per-rank counter
vectors (good / bad-errors / bad-length / bad-ambigs / near-cutoff / ... + 64-bin floor(ee)
histogram, MOIRA_N_COUNTERS x uint64) are summed with one all-reduce -- NCCL over NVLink on GPUs,
gloo in the CPU tests.  torch.distributed is plumbing only; no read data crosses ranks.
"""
from __future__ import annotations

import numpy as np


def shard_range(n_reads: int, rank: int, world: int):
    """[begin, end) of rank's contiguous chunk."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    return n_reads * rank // world, n_reads * (rank + 1) // world


def shard_slab(offsets, lengths, rank: int, world: int):
    """Slice packed-slab metadata for one rank; returns (begin, end, byte_begin, byte_end)."""
    n = len(lengths)
    b, e = shard_range(n, rank, world)
    if b == e:
        return b, e, 0, 0
    byte_b = int(offsets[b])
    byte_e = int(offsets[e - 1]) + (int(lengths[e - 1]) + 15) // 16 * 16
    return b, e, byte_b, byte_e


def reduce_counters(counters, group=None):
    """Sum a MOIRA_N_COUNTERS vector over all ranks.  `counters` is a torch int64 tensor (on the GPU
    for NCCL, on the CPU for gloo) or a numpy uint64 array (copied through a CPU tensor)."""
    import torch
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return counters
    if isinstance(counters, np.ndarray):
        t = torch.from_numpy(counters.astype(np.int64))
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        return t.numpy().astype(np.uint64)
    dist.all_reduce(counters, op=dist.ReduceOp.SUM, group=group)
    return counters
