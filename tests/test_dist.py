"""CPU: the N>1 host logic -- contiguous sharding and the counter all-reduce -- with two gloo ranks.
The per-shard counters are produced by the oracle here (no GPU in this tier); on the B200 box the
same reduce runs over NCCL inside bench.py."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from moira_b200 import shard, synth  # noqa: E402
from moira_b200 import _lib as L  # noqa: E402
from oracle import py_oracle as po  # noqa: E402


def _counters(slab, off, ln):
    ee, ns = po.pb_batch(slab, off, ln, 0.005)
    ok, reason, eef = po.decide_batch(ee, ns, ln, np.zeros(len(ln), bool), thr_kind="uncert", thr=0.01,
                                      ambigs="treat_as_errors", round_flag=False, truncate=None)
    c = np.zeros(L.N_COUNTERS, np.uint64)
    c[L.CNT_READS] = len(ln)
    c[L.CNT_ACCEPTED] = int(ok.sum())
    c[L.CNT_BAD_ERRORS] = int((~ok).sum())
    c[L.CNT_HIST:L.CNT_HIST + 64] = np.bincount(np.minimum(np.floor(eef), 63).astype(int), minlength=64)
    return c


def _worker(rank, world, port, n, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    slab, off, ln = synth.generate("v4", n, 99)
    b, e, _, _ = shard.shard_slab(off, ln, rank, world)
    local = _counters(slab, off[b:e], ln[b:e])
    total = shard.reduce_counters(local)
    t = torch.from_numpy(local.astype(np.int64))
    total_t = shard.reduce_counters(t.clone())
    assert np.array_equal(total_t.numpy().astype(np.uint64), total)
    np.save(os.path.join(out_dir, "rank%d.npy" % rank), total)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_cover_exactly_once():
    for n in (0, 1, 7, 1000, 12345):
        for world in (1, 2, 3, 8):
            spans = [shard.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard.shard_range(10, 2, 2)


def test_two_rank_gloo_counter_reduce(tmp_path):
    n, world = 3000, 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(world, port, n, str(tmp_path)), nprocs=world, join=True)
    slab, off, ln = synth.generate("v4", n, 99)
    whole = _counters(slab, off, ln)
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / ("rank%d.npy" % r)), whole)


def test_read_pairs_shard_like_reads(oracle_contigs, forward_records, reverse_records):
    """Pairs are as independent as reads: rank r of N builds the contigs of its contiguous range of pairs and the
    concatenation over ranks is the single-rank result (what `moira_filter_pairs` per rank + one counter reduce gives)."""
    n = 120
    whole = [c[1:] for c in oracle_contigs[:n]]
    for world in (2, 3):
        parts = []
        for r in range(world):
            b, e = shard.shard_range(n, r, world)
            parts += [po.pair_to_contig(forward_records[i][1], forward_records[i][2], reverse_records[i][1], reverse_records[i][2])
                      for i in range(b, e)]
        assert parts == whole
