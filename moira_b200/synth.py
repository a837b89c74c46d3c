"""Seeded synthetic reads in slab form, for the parity tests and bench.py.

Profiles (calibrated in SURVEY.md 8d against the reference's fixtures under moira/test/):
  v4      253-bp 16S V4 contigs, MiSeq: 80 % clean / 20 % noisy reads, i.i.d. Q per class,
          first 10 bases 8 lower; a base drawn with Q == 2 becomes 'N' with probability 0.5
  v3v4    ~450-bp (420..480) contigs, quality dips in the middle of the contig
  ccs     1500-bp CCS-like full-length 16S, uint8-safe Q <= 93
  mixed   lengths uniform in 100..600 with the v3v4 shape
  real    253-bp rows bootstrapped from the reference's real contig fixture with +-1 jitter (8d's preferred generator)

`generate(profile, n, seed)` is the numpy (host) generator used where the CPU oracle has to see the
same bytes; `generate_v4_device` builds the same distribution with torch on the GPU (different
random stream) for bench.py, whose 10 M-read slab would take minutes on the host.
Slab rows are padded to a multiple of 16 bytes with 0xFD and start 16-byte aligned.
"""
from __future__ import annotations

import numpy as np

N_MARK, PAD = 0xFF, 0xFD

# class histograms for the v4 profile: (Q values, probabilities)
_CLEAN_Q = np.array([38, 39, 40, 37, 36, 35, 34, 33, 32, 31, 30], dtype=np.int64)
_CLEAN_P = np.array([.30, .32, .12, .06, .07, .05, .016, .016, .016, .016, .016])
_NOISY_Q = np.array([2] + list(range(6, 36)) + [36, 37, 38, 39, 40], dtype=np.int64)
_NOISY_P = np.array([.064] + [.02] * 30 + [.05, .06, .09, .09, .046])
_CCS_Q = np.array([93] + list(range(60, 93)) + list(range(30, 60)) + list(range(10, 30)) + list(range(2, 10)),
                  dtype=np.int64)
_CCS_P = np.array([.55] + [.25 / 33] * 33 + [.15 / 30] * 30 + [.04 / 20] * 20 + [.01 / 8] * 8)
V4_LEN = 253
V4_STRIDE = 256
V4_NOISY_FRACTION = 0.20


def _norm(p):
    p = np.asarray(p, dtype=np.float64)
    return p / p.sum()


def v4_tables():
    """(clean_q, clean_cdf, noisy_q, noisy_cdf) shared by the host and device generators."""
    return _CLEAN_Q, np.cumsum(_norm(_CLEAN_P)), _NOISY_Q, np.cumsum(_norm(_NOISY_P))


def _rows_to_slab(rows_q, lengths):
    """rows_q: int array [n, Lmax] of slab byte values; lengths uint32[n] -> (slab, offsets, lengths)."""
    n, lmax = rows_q.shape
    stride = (lmax + 15) // 16 * 16
    slab = np.full((n, stride), PAD, dtype=np.uint8)
    slab[:, :lmax] = rows_q.astype(np.uint8)
    cols = np.arange(stride)[None, :]
    slab[cols >= lengths[:, None]] = PAD
    offsets = (np.arange(n, dtype=np.uint64) * np.uint64(stride))
    return slab.reshape(-1), offsets, lengths.astype(np.uint32)


def _apply_n(rng, q):
    is2 = (q == 2) & (rng.random(q.shape) < 0.5)          # moira.py:1496-1497, 1532-1533: introduced Ns get Q=2
    return np.where(is2, N_MARK, q)


def generate(profile: str, n: int, seed: int):
    """Returns (slab uint8, offsets uint64, lengths uint32).  Fixed-length profiles have a uniform stride."""
    rng = np.random.Generator(np.random.PCG64(seed))
    if profile == "v4":
        cq, ccdf, nq, ncdf = v4_tables()
        noisy = rng.random(n) < V4_NOISY_FRACTION
        u = rng.random((n, V4_LEN))
        qc = cq[np.minimum(np.searchsorted(ccdf, u), len(cq) - 1)]
        qn = nq[np.minimum(np.searchsorted(ncdf, u), len(nq) - 1)]
        q = np.where(noisy[:, None], qn, qc)
        q[:, :10] = np.maximum(q[:, :10] - 8, 2)
        q = _apply_n(rng, q)
        return _rows_to_slab(q, np.full(n, V4_LEN, dtype=np.uint32))
    if profile in ("v3v4", "mixed"):
        if profile == "v3v4":
            lengths = np.clip(np.rint(rng.normal(450, 8, n)), 420, 480).astype(np.int64)
        else:
            lengths = rng.integers(100, 601, n)
        lmax = int(lengths.max())
        i = np.arange(lmax)[None, :]
        L = lengths[:, None].astype(np.float64)
        mu = 38 - 12 * (1 - np.abs(2 * i / L - 1)) ** 2
        d = rng.gamma(2.0, 1.0, n)[:, None]
        q = np.clip(np.rint(mu - rng.exponential(1.0, (n, lmax)) * d), 2, 40).astype(np.int64)
        q = _apply_n(rng, q)
        return _rows_to_slab(q, lengths.astype(np.uint32))
    if profile == "ccs":
        L = 1500
        cdf = np.cumsum(_norm(_CCS_P))
        # per-read accuracy class: most reads near-perfect, some degraded
        shift = rng.choice([0, 0, 0, 10, 25], size=n)[:, None]
        q = _CCS_Q[np.minimum(np.searchsorted(cdf, rng.random((n, L))), len(_CCS_Q) - 1)]
        q = np.clip(q - shift * (rng.random((n, L)) < 0.5), 2, 93)
        q = _apply_n(rng, q)
        return _rows_to_slab(q, np.full(n, L, dtype=np.uint32))
    if profile == "real":
        # SURVEY.md 8d's preferred generator: bootstrap whole quality rows of the reference's real 253-bp contig
        # fixture (test_results/paired.qc.*, staged in tests/golden/contigs.json.gz) with +-1 jitter
        rows_q, rows_n = real_rows()
        idx = rng.integers(0, rows_q.shape[0], n)
        q = np.clip(rows_q[idx].astype(np.int64) + rng.integers(-1, 2, (n, rows_q.shape[1])), 1, 41)
        q = np.where(rows_n[idx], N_MARK, q)
        return _rows_to_slab(q, np.full(n, rows_q.shape[1], dtype=np.uint32))
    raise ValueError("unknown profile %r" % profile)


_REAL = None


def real_rows():
    """(qualities uint8 [328, 253], is-N bool [328, 253]) of the length-253 contigs of the reference's paired fixture."""
    global _REAL
    if _REAL is None:
        import gzip
        import json
        import os
        path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "contigs.json.gz")
        with gzip.open(path) as fh:
            recs = [c for c in json.load(fh) if len(c["seq"]) == V4_LEN]
        q = np.array([c["quals"] for c in recs], dtype=np.uint8)
        isn = np.array([[ch in "Nn" for ch in c["seq"]] for c in recs], dtype=bool)
        _REAL = (q, isn)
    return _REAL


# (stride, fixed length or None) of the device generators' slabs
DEVICE_LAYOUT = {"v4": (256, 253), "real": (256, 253), "v3v4": (480, None), "mixed": (608, None), "ccs": (1504, 1500)}


def generate_device(profile: str, n: int, seed: int, device, chunk: int = 1 << 19, with_sequences: bool = False):
    """torch (device-side) generators of the profiles above (same distributions, different random stream).
    Returns (slab uint8 [n, stride], lengths int32 [n] or None when every read has DEVICE_LAYOUT[profile][1] bases,
    sequences uint8 [n, stride] or None).  Sequences (ccs, for --collapse): a pool of 250 000 distinct strings with
    Zipf(1.2) abundance, 'N' where the slab has the marker."""
    import torch

    stride, fixed = DEVICE_LAYOUT[profile]
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    if profile == "v4":
        return generate_v4_device(n, seed, device), None, None
    out = torch.full((n, stride), PAD, dtype=torch.uint8, device=device)
    lengths = None
    seqs = None
    if profile == "real":
        rq, rn = real_rows()
        rq_t = torch.tensor(rq, device=device, dtype=torch.int16)
        rn_t = torch.tensor(rn, device=device)
        for s0 in range(0, n, chunk):
            m = min(chunk, n - s0)
            idx = torch.randint(0, rq.shape[0], (m,), generator=g, device=device)
            jit = torch.randint(-1, 2, (m, V4_LEN), generator=g, device=device, dtype=torch.int16)
            q = (rq_t[idx] + jit).clamp_(1, 41).to(torch.uint8)
            q = torch.where(rn_t[idx], torch.full_like(q, N_MARK), q)
            out[s0:s0 + m, :V4_LEN] = q
        return out, None, None
    if profile in ("v3v4", "mixed"):
        lengths = torch.empty(n, dtype=torch.int32, device=device)
        lmax = 480 if profile == "v3v4" else 600
        pos = torch.arange(lmax, device=device, dtype=torch.float32)[None, :]
        for s0 in range(0, n, chunk):
            m = min(chunk, n - s0)
            if profile == "v3v4":
                ln = torch.randn(m, generator=g, device=device).mul_(8).add_(450).round_().clamp_(420, 480)
            else:
                ln = torch.randint(100, 601, (m,), generator=g, device=device).to(torch.float32)
            lengths[s0:s0 + m] = ln.to(torch.int32)
            L = ln[:, None]
            mu = 38 - 12 * (1 - (2 * pos / L - 1).abs()) ** 2
            d = -(torch.rand((m, 1), generator=g, device=device).clamp_(min=1e-12).log()
                  + torch.rand((m, 1), generator=g, device=device).clamp_(min=1e-12).log())            # Gamma(2, 1)
            e = -torch.rand((m, lmax), generator=g, device=device).clamp_(min=1e-12).log()          # Exp(1)
            q = (mu - e * d).round_().clamp_(2, 40).to(torch.uint8)
            isn = (q == 2) & (torch.rand((m, lmax), generator=g, device=device) < 0.5)
            q = torch.where(isn, torch.full_like(q, N_MARK), q)
            q = torch.where(pos < L, q, torch.full_like(q, PAD))
            out[s0:s0 + m, :lmax] = q
            del mu, e, q, isn
        return out, lengths, None
    if profile == "ccs":
        L = 1500
        cdf = torch.tensor(np.cumsum(_norm(_CCS_P)), device=device, dtype=torch.float32)
        qv = torch.tensor(_CCS_Q, device=device, dtype=torch.int16)
        shifts = torch.tensor([0, 0, 0, 10, 25], device=device, dtype=torch.int16)
        pool = pool_cdf = None
        if with_sequences:
            n_pool = 250_000
            pool = torch.tensor([65, 67, 71, 84], device=device, dtype=torch.uint8)[
                torch.randint(0, 4, (n_pool, L), generator=g, device=device)]
            w = torch.arange(1, n_pool + 1, device=device, dtype=torch.float64).pow_(-1.2)
            pool_cdf = (w.cumsum(0) / w.sum()).to(torch.float32)
            seqs = torch.zeros((n, stride), dtype=torch.uint8, device=device)
        sub = max(1, chunk // 4)
        for s0 in range(0, n, sub):
            m = min(sub, n - s0)
            q = qv[torch.searchsorted(cdf, torch.rand((m, L), generator=g, device=device)).clamp_(max=len(_CCS_Q) - 1)]
            sh = shifts[torch.randint(0, 5, (m, 1), generator=g, device=device)]
            q = (q - sh * (torch.rand((m, L), generator=g, device=device) < 0.5)).clamp_(2, 93).to(torch.uint8)
            isn = (q == 2) & (torch.rand((m, L), generator=g, device=device) < 0.5)
            q = torch.where(isn, torch.full_like(q, N_MARK), q)
            out[s0:s0 + m, :L] = q
            if with_sequences:
                pid = torch.searchsorted(pool_cdf, torch.rand(m, generator=g, device=device)).clamp_(max=pool.shape[0] - 1)
                sq = pool[pid]
                seqs[s0:s0 + m, :L] = torch.where(isn, torch.full_like(sq, 78), sq)
            del q, isn
        return out, None, seqs
    raise ValueError("unknown profile %r" % profile)


def generate_v4_device(n: int, seed: int, device, chunk: int = 1 << 20):
    """torch (device-side) generator of the v4 profile: returns a uint8 tensor [n, 256] on `device`."""
    import torch

    cq, ccdf, nq, ncdf = v4_tables()
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    cq_t = torch.tensor(cq, device=device, dtype=torch.uint8)
    nq_t = torch.tensor(nq, device=device, dtype=torch.uint8)
    ccdf_t = torch.tensor(ccdf, device=device, dtype=torch.float32)
    ncdf_t = torch.tensor(ncdf, device=device, dtype=torch.float32)
    out = torch.full((n, V4_STRIDE), PAD, dtype=torch.uint8, device=device)
    for s in range(0, n, chunk):
        m = min(chunk, n - s)
        u = torch.rand((m, V4_LEN), generator=g, device=device)
        noisy = torch.rand((m, 1), generator=g, device=device) < V4_NOISY_FRACTION
        ic = torch.searchsorted(ccdf_t, u).clamp_(max=len(cq) - 1)
        inn = torch.searchsorted(ncdf_t, u).clamp_(max=len(nq) - 1)
        q = torch.where(noisy, nq_t[inn], cq_t[ic])
        head = q[:, :10].to(torch.int16) - 8
        q[:, :10] = head.clamp_(min=2).to(torch.uint8)
        isn = (q == 2) & (torch.rand((m, V4_LEN), generator=g, device=device) < 0.5)
        q = torch.where(isn, torch.full_like(q, N_MARK), q)
        out[s:s + m, :V4_LEN] = q
        del u, ic, inn, q, isn
    return out


# algorithmic work per read, SURVEY.md 8(d)
def w_fp64(length: int, k: int) -> int:
    return (length - 1) * (3 * k - 2) + (k - 1) + 6


def w_hbm(length: int) -> int:
    return (length + 15) // 16 * 16 + 16


def decision_k(length: int, uncert: float = 0.01) -> int:
    import math
    return int(math.floor(length * uncert)) + 2
