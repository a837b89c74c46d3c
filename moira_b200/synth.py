"""Seeded synthetic reads in slab form, for the parity tests and bench.py.

Profiles (calibrated in SURVEY.md 8d against the reference's fixtures under moira/test/):
  v4      253-bp 16S V4 contigs, MiSeq: 80 % clean / 20 % noisy reads, i.i.d. Q per class,
          first 10 bases 8 lower; a base drawn with Q == 2 becomes 'N' with probability 0.5
  v3v4    ~450-bp (420..480) contigs, quality dips in the middle of the contig
  ccs     1500-bp CCS-like full-length 16S, uint8-safe Q <= 93
  mixed   lengths uniform in 100..600 with the v3v4 shape

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
