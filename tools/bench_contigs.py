#!/usr/bin/env python3
"""pairs/s of the paired-end contig constructor on one B200, with the reference timed beside it.

  python tools/bench_contigs.py [--pairs N] [--read-len L] [--steps K] > profiles/r01_contigs.json

Workload: N synthetic MiSeq V4 pairs (2 x L bases sequenced from a 253-bp amplicon, so the reads overlap almost
completely, like moira/test/test{1,2}.fastq; substitution errors where the quality is low; seed 20160401).
  kernel   contig kernel only (CUDA events around its launches, inputs on the device)
  e2e      moira_filter_pairs with host arrays: H2D of both reads, contig kernel, filter kernels, D2H of contigs,
           statistics and decisions
  cpu      the unmodified reference aligner (oracle/_ref/nw_align.so, cythonised nw_align.pyx) on one core, and the
           whole pair -> contig step of the C restatement on all cores (multiprocessing)
Algorithmic work of a pair: L1 x L2 matrix cells (nw_align.pyx:87-116); the kernel is integer-issue bound.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

COMP = str.maketrans("ACGTN", "TGCAN")


def make_pairs(n, read_len, seed=20160401, amplicon=253):
    rng = np.random.Generator(np.random.PCG64(seed))
    frag = rng.integers(0, 4, size=(n, max(amplicon, read_len)), dtype=np.uint8)
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    qprof = np.clip(np.round(38 - 28 * (np.arange(read_len) / read_len) ** 3), 2, 38).astype(np.int64)   # decays along the read
    out = []
    for strand in (0, 1):
        src = frag[:, :read_len] if strand == 0 else (3 - frag[:, ::-1][:, :read_len])       # reverse complement in 0..3 code
        q = np.clip(qprof[None, :] - rng.geometric(0.35, size=(n, read_len)) + 1, 2, 40)
        err = rng.random((n, read_len)) < 10.0 ** (-q / 10.0)
        bases = np.where(err, (src + rng.integers(1, 4, size=src.shape, dtype=np.uint8)) & 3, src)
        out.append((lut[bases].reshape(-1).copy(), q.astype(np.uint8).reshape(-1).copy(),
                    np.arange(n, dtype=np.uint64) * read_len, np.full(n, read_len, np.uint32)))
    return out


def _pool_pairs(chunk):
    from oracle import py_oracle as po
    return [po.pair_to_contig(*p)[2] for p in chunk]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=1 << 19)
    ap.add_argument("--read-len", type=int, default=251)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    pool = None
    if not args.no_cpu:
        import multiprocessing as mp
        pool = mp.get_context("fork").Pool(os.cpu_count() or 1)     # before CUDA is initialised

    import moira_b200
    from moira_b200 import ContigParams, FilterParams
    from moira_b200 import _lib as L
    from oracle import py_oracle as po

    n, rl = args.pairs, args.read_len
    fwd, rev = make_pairs(n, rl)
    ctx = moira_b200.Context(0)
    cp, fp = ContigParams(), FilterParams(exact_ee=False)
    # page-locked host buffers on both sides, reused by every step
    pins = []

    def pinned(arr):
        pb = moira_b200.PinnedBuffer(arr.nbytes)
        pins.append(pb)
        v = pb.view(arr.dtype, arr.size)
        v[:] = arr
        return v

    fwd = (pinned(fwd[0]), pinned(fwd[1]), fwd[2], fwd[3])
    rev = (pinned(rev[0]), pinned(rev[1]), rev[2], rev[3])
    stride = (2 * rl + 15) // 16 * 16
    out_buf = moira_b200.PairResult.allocate(n, stride, True, pinned=True)
    res = ctx.filter_pairs(*fwd, *rev, cp, fp, out=out_buf)          # warm-up (allocations, tables)
    assert not res.status.any()
    # parity sample against the oracle
    bad = 0
    for r in list(range(0, 200)) + list(range(n - 200, n)):
        f = (fwd[0][r * rl:(r + 1) * rl].tobytes().decode(), fwd[1][r * rl:(r + 1) * rl].tolist())
        v = (rev[0][r * rl:(r + 1) * rl].tobytes().decode(), rev[1][r * rl:(r + 1) * rl].tolist())
        want = po.pair_to_contig(f[0], f[1], v[0], v[1])
        bad += res.contig(r) != (want[0], want[1]) or (int(res.overlap[r]), int(res.gaps[r]), int(res.mismatches[r])) != want[2:]
    ctx.set_timing(True)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = ctx.filter_pairs(*fwd, *rev, cp, fp, out=out_buf)
    dt = (time.perf_counter() - t0) / args.steps
    kms, launches = ctx.last_contig_ms()
    fms, _ = ctx.last_kernel_ms()
    ctx.set_timing(False)
    kms /= args.steps
    cells = float(n) * rl * rl
    out = {
        "metric": "pairs/s paired-end contig construction (2 x %d bp)" % rl, "unit": "pairs/s", "n_gpus": 1, "steps": args.steps,
        "config": {"workload": "%d synthetic MiSeq V4 pairs, 2 x %d bp over a 253-bp amplicon, match 1 / mismatch -1 / gap -2, "
                               "consensus best, cap 40; then the Poisson-binomial filter on the contigs" % (n, rl)},
        "kernel": {"value": n / (kms * 1e-3), "ms_per_step": kms, "launches_per_step": launches / args.steps,
                   "cells_per_s": cells / (kms * 1e-3), "note": "contig kernel only, CUDA events on its stream"},
        "e2e": {"value": n / dt, "ms_per_step": dt * 1e3, "filter_first_pass_ms": fms / args.steps,
                "h2d_bytes_per_step": int(sum(a.nbytes for a in fwd) + sum(a.nbytes for a in rev)),
                "d2h_bytes_per_step": int(res.contig_seq.nbytes + res.contig_qual.nbytes + n * (4 * 4 + 1 + 8 + 4 + 1)),
                "api": "moira_filter_pairs (pinned host arrays in, contigs + statistics + decisions out)"},
        "accepted_fraction": float(res.filter.counters[L.CNT_ACCEPTED]) / n,
        "mean_contig_len": float(res.contig_len.mean()),
        "parity": {"sample_pairs": 400, "mismatches": int(bad), "checker": "oracle/contig_oracle.c"},
    }
    if pool is not None:
        cpu = {}
        if po.have_ref_nw():
            nw = po.ref_nw_module()
            k = 300
            pairs = [(fwd[0][r * rl:(r + 1) * rl].tobytes().decode(), po.reverse_complement(rev[0][r * rl:(r + 1) * rl].tobytes().decode()))
                     for r in range(k)]
            t0 = time.perf_counter()
            for a, b in pairs:
                nw.nw_align(a, b, 1, -1, -2)
            cpu["reference_nw_align_1core"] = {"value": k / (time.perf_counter() - t0), "unit": "pairs/s", "kind": "reference",
                                               "sample": "%d pairs through the unmodified nw_align.pyx (alignment only), 1 core" % k}
        cores = os.cpu_count() or 1
        k = 400 * cores
        items = [(fwd[0][r * rl:(r + 1) * rl].tobytes().decode(), fwd[1][r * rl:(r + 1) * rl].tolist(),
                  rev[0][r * rl:(r + 1) * rl].tobytes().decode(), rev[1][r * rl:(r + 1) * rl].tolist()) for r in range(k)]
        chunks = [items[i:i + 100] for i in range(0, k, 100)]
        pool.map(_pool_pairs, chunks[:cores])
        t0 = time.perf_counter()
        pool.map(_pool_pairs, chunks, chunksize=1)
        cpu["oracle_pair_to_contig_allcores"] = {"value": k / (time.perf_counter() - t0), "unit": "pairs/s", "kind": "port", "cores": cores,
                                                 "sample": "%d pairs, reverse complement + alignment + consensus in C (oracle), %d processes" % (k, cores)}
        pool.terminate()
        out["cpu_baseline"] = cpu
    print(json.dumps(out))
    ctx.close()


if __name__ == "__main__":
    main()
