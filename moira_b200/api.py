"""Batched host API over libmoira_b200.so -- the call a host program makes instead of looping
`bernoulli.calculate_errors_PB` (reference: moira/moira.py:815-831 per read, moira.py:925-970 for
the decision).  Everything numeric happens in the CUDA library; this module only marshals buffers.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass

import numpy as np

from . import _lib as L
from ._lib import MoiraError, lib

_MODES = {"poisson_binomial": L.MODE_PB, "poisson": L.MODE_POISSON, "expected_error": L.MODE_EXPECTED_ERROR}
_AMBIGS = {"treat_as_errors": L.AMBIGS_TREAT_AS_ERRORS, "ignore": L.AMBIGS_IGNORE, "disallow": L.AMBIGS_DISALLOW}


@dataclass
class FilterParams:
    """Hot-path relevant moira flags, with the reference's names and defaults (moira.py:649-668)."""
    error_calc: str = "poisson_binomial"   # --error_calc (+ north_star's added 'expected_error')
    alpha: float = 0.005                   # --alpha
    uncert: float | None = 0.01            # --uncert
    maxerrors: float | None = None         # --maxerrors (used when truthy, like `if args.maxerrors:` moira.py:925)
    ambigs: str = "treat_as_errors"        # --ambigs
    round: bool = False                    # --round
    truncate: int | None = None            # --truncate
    exact_ee: bool = True                  # exact statistic for every read (needed by --collapse, moira.py:466)
    ee_output: str = "raw"                 # 'raw' = calculate_errors_* value, 'final' = process_data value
    slab_format: str = "q8"                # host slab format: 'q8' (one byte per base) or 'q6' (pack_q6 transport image)
    length_sort: int = 0                   # ragged batches: 0 = bucket by length on the device when it pays, 2 = never

    def to_c(self) -> L.Params:
        if self.error_calc not in _MODES:
            raise ValueError("error_calc must be one of %s" % sorted(_MODES))
        if self.ambigs not in _AMBIGS:
            raise ValueError("ambigs must be one of %s" % sorted(_AMBIGS))
        p = L.Params()
        p.mode = _MODES[self.error_calc]
        if self.maxerrors:
            p.thr_kind, p.thr = L.THR_MAXERRORS, float(self.maxerrors)
        else:
            p.thr_kind, p.thr = L.THR_UNCERT, float(self.uncert if self.uncert is not None else 0.01)
        p.ambigs = _AMBIGS[self.ambigs]
        p.round_flag = 1 if self.round else 0
        p.truncate = int(self.truncate) if self.truncate else 0
        p.exact_ee = 1 if self.exact_ee else 0
        p.ee_output = L.EE_FINAL if self.ee_output == "final" else L.EE_RAW
        p.length_sort = int(self.length_sort)
        p.slab_format = L.SLAB_Q6 if self.slab_format == "q6" else L.SLAB_Q8
        p.alpha = float(self.alpha)
        return p

    @property
    def lower_n_ambiguous(self) -> bool:
        """The C core skips 'N' and 'n' (bernoullimodule.c:196); the Python calculators only 'N'
        (moira.py:1605, 1660)."""
        return self.error_calc == "poisson_binomial"


@dataclass
class FilterResult:
    ee: np.ndarray        # float64[n]
    ns: np.ndarray        # int32[n]
    flags: np.ndarray     # uint8[n]
    counters: np.ndarray  # uint64[N_COUNTERS]

    @property
    def accept(self) -> np.ndarray:
        return (self.flags & L.FLAG_ACCEPT) != 0

    @property
    def reason(self) -> np.ndarray:
        return (self.flags & L.FLAG_REASON_MASK) >> 1

    @property
    def lower_bound(self) -> np.ndarray:
        return (self.flags & L.FLAG_LOWER_BOUND) != 0

    @property
    def near_cutoff(self) -> np.ndarray:
        return (self.flags & L.FLAG_NEAR_CUTOFF) != 0

    @property
    def numeric(self) -> np.ndarray:
        return (self.flags & L.FLAG_NUMERIC) != 0


class PinnedBuffer:
    """Page-locked host memory (moira_host_alloc) exposed as a numpy array."""

    def __init__(self, nbytes: int):
        self._ptr = ctypes.c_void_p()
        L.check(lib.moira_host_alloc(ctypes.byref(self._ptr), int(nbytes)))
        self.nbytes = int(nbytes)
        buf = (ctypes.c_uint8 * max(1, self.nbytes)).from_address(self._ptr.value)
        self.u8 = np.frombuffer(buf, dtype=np.uint8, count=self.nbytes)

    def view(self, dtype, count=None, offset=0):
        dt = np.dtype(dtype)
        count = (self.nbytes - offset) // dt.itemsize if count is None else count
        return self.u8[offset:offset + count * dt.itemsize].view(dt)

    def free(self):
        if self._ptr:
            self.u8 = None
            lib.moira_host_free(self._ptr)
            self._ptr = ctypes.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def _ptr(a):
    return None if a is None else a.ctypes.data


def _as(a, dtype):
    a = np.ascontiguousarray(a, dtype=dtype)
    return a


class Context:
    """One moira_ctx: one GPU, its streams, tables and workspaces.  Not thread-safe."""

    def __init__(self, device: int = 0):
        self._h = ctypes.c_void_p()
        L.check(lib.moira_ctx_create(int(device), ctypes.byref(self._h)))
        self.device = int(device)

    def close(self):
        if self._h:
            lib.moira_ctx_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- info ---------------------------------------------------------------------------------
    @property
    def sm_count(self) -> int:
        v = ctypes.c_int()
        L.check(lib.moira_ctx_sm_count(self._h, ctypes.byref(v)))
        return v.value

    def lut(self):
        p, q, e = (np.zeros(256), np.zeros(256), np.zeros(256))
        L.check(lib.moira_ctx_get_lut(self._h, p.ctypes.data, q.ctypes.data, e.ctypes.data))
        return p, q, e

    @property
    def launch_count(self) -> int:
        v = ctypes.c_uint64()
        L.check(lib.moira_ctx_launch_count(self._h, ctypes.byref(v)))
        return v.value

    def set_timing(self, enabled: bool):
        L.check(lib.moira_ctx_set_timing(self._h, 1 if enabled else 0))

    def last_kernel_ms(self):
        ms = ctypes.c_float()
        name = ctypes.c_char_p()
        L.check(lib.moira_ctx_last_kernel_ms(self._h, ctypes.byref(ms), ctypes.byref(name)))
        return ms.value, (name.value or b"").decode()

    def fp64_peak(self, iters: int = 20000):
        ops = ctypes.c_double()
        ms = ctypes.c_double()
        L.check(lib.moira_fp64_peak(self._h, int(iters), ctypes.byref(ops), ctypes.byref(ms)))
        return ops.value, ms.value

    # ---- hot path -------------------------------------------------------------------------------
    def filter_batch(self, slab, offsets, lengths, params: FilterParams, out: FilterResult | None = None) -> FilterResult:
        """Host buffers in, host buffers out (moira_filter_batch)."""
        slab = _as(slab, np.uint8)
        offsets = _as(offsets, np.uint64)
        lengths = _as(lengths, np.uint32)
        n = int(lengths.shape[0])
        if out is None:
            out = FilterResult(np.empty(n, np.float64), np.empty(n, np.int32), np.empty(n, np.uint8),
                               np.zeros(L.N_COUNTERS, np.uint64))
        cp = params.to_c()
        L.check(lib.moira_filter_batch(self._h, _ptr(slab), slab.nbytes, _ptr(offsets), _ptr(lengths), n,
                                       ctypes.byref(cp), _ptr(out.ee), _ptr(out.ns), _ptr(out.flags),
                                       _ptr(out.counters)))
        return out

    def submit(self, slab, offsets, lengths, params: FilterParams, out: FilterResult) -> int:
        """Asynchronous moira_submit; all arrays must stay alive until wait(ticket)."""
        n = int(lengths.shape[0])
        cp = params.to_c()
        t = ctypes.c_int()
        L.check(lib.moira_submit(self._h, _ptr(slab), slab.nbytes, _ptr(offsets), _ptr(lengths), n, ctypes.byref(cp),
                                 _ptr(out.ee), _ptr(out.ns), _ptr(out.flags), _ptr(out.counters), ctypes.byref(t)))
        return t.value

    def wait(self, ticket: int):
        L.check(lib.moira_wait(self._h, int(ticket)))

    def filter_device(self, d_slab: int, d_offsets: int | None, d_lengths: int | None, stride: int,
                      fixed_length: int, n_reads: int, params: FilterParams, d_ee: int, d_ns: int | None,
                      d_flags: int | None, d_counters: int | None, stream: int | None = None):
        """Device pointers (ints, e.g. torch.Tensor.data_ptr()); enqueues on `stream`, no sync."""
        cp = params.to_c()
        L.check(lib.moira_filter_device(self._h, d_slab, d_offsets, d_lengths, int(stride), int(fixed_length),
                                        int(n_reads), ctypes.byref(cp), d_ee, d_ns, d_flags, d_counters, stream))

    def filter_fastq(self, text: bytes, params: FilterParams, fastq_offset: int = 33, out: FilterResult | None = None):
        """FASTQ text -> FilterResult in one streaming C call (moira_filter_fastq); also returns lengths."""
        buf = np.frombuffer(text, dtype=np.uint8)
        n = ctypes.c_uint64()
        L.check(lib.moira_fastq_count_reads(_ptr(buf), buf.nbytes, ctypes.byref(n)))
        nr = n.value
        if out is None:
            out = FilterResult(np.empty(nr, np.float64), np.empty(nr, np.int32), np.empty(nr, np.uint8),
                               np.zeros(L.N_COUNTERS, np.uint64))
        lengths = np.empty(nr, np.uint32)
        cp = params.to_c()
        L.check(lib.moira_filter_fastq(self._h, _ptr(buf), buf.nbytes, int(fastq_offset), int(params.lower_n_ambiguous),
                                       ctypes.byref(cp), nr, _ptr(out.ee), _ptr(out.ns), _ptr(out.flags), _ptr(lengths),
                                       _ptr(out.counters), ctypes.byref(n)))
        return out, lengths

    def calculate_errors_PB(self, contig: str, contig_quals, alpha: float):
        """One read through the CUDA path (moira_calculate_errors_PB)."""
        q = np.ascontiguousarray(contig_quals, dtype=np.int32)
        ee = ctypes.c_double()
        ns = ctypes.c_int32()
        rc = lib.moira_calculate_errors_PB(self._h, contig.encode("latin-1"), _ptr(q) if q.size else None,
                                           int(q.shape[0]), float(alpha), ctypes.byref(ee), ctypes.byref(ns))
        L.check(rc)
        return ee.value, ns.value


# ---- host-side packing (C++ in the library, no GPU) ---------------------------------------------
def build_lut():
    p, q, e = np.zeros(256), np.zeros(256), np.zeros(256)
    eqp = ctypes.c_int()
    L.check(lib.moira_build_lut(p.ctypes.data, q.ctypes.data, e.ctypes.data, ctypes.byref(eqp)))
    return p, q, e, bool(eqp.value)


def pack_reads(seqs, quals_list, lower_n_ambiguous: bool = True):
    """Pack python reads (str, list[int]) into the in-band slab via moira_pack_reads.
    Returns (slab uint8, offsets uint64, lengths uint32)."""
    n = len(seqs)
    lengths = np.fromiter((len(s) for s in seqs), dtype=np.uint32, count=n)
    in_off = np.zeros(n, dtype=np.uint64)
    if n:
        in_off[1:] = np.cumsum(lengths[:-1], dtype=np.uint64)
    seq_all = np.frombuffer("".join(seqs).encode("latin-1"), dtype=np.uint8) if n else np.zeros(0, np.uint8)
    total = int(lengths.sum()) if n else 0
    q_all = np.empty(total, dtype=np.int32)
    pos = 0
    for s, q in zip(seqs, quals_list):
        if len(q) != len(s):
            raise ValueError("contig and contig_quals must have the same length")
        q_all[pos:pos + len(q)] = q
        pos += len(q)
    return pack_arrays(seq_all, q_all, in_off, lengths, lower_n_ambiguous)


def pack_arrays(seq_all, q_all, in_off, lengths, lower_n_ambiguous: bool = True, slab_out=None):
    seq_all = _as(seq_all, np.uint8)
    q_all = _as(q_all, np.int32)
    in_off = _as(in_off, np.uint64)
    lengths = _as(lengths, np.uint32)
    n = int(lengths.shape[0])
    need = ctypes.c_uint64()
    L.check(lib.moira_pack_reads(_ptr(seq_all), _ptr(q_all), _ptr(in_off), _ptr(lengths), n,
                                 int(lower_n_ambiguous), None, 0, None, ctypes.byref(need)))
    nbytes = max(16, need.value)
    slab = slab_out if slab_out is not None else np.empty(nbytes, dtype=np.uint8)
    if slab.nbytes < nbytes:
        raise ValueError("slab_out too small")
    offsets = np.empty(n, dtype=np.uint64)
    L.check(lib.moira_pack_reads(_ptr(seq_all), _ptr(q_all), _ptr(in_off), _ptr(lengths), n,
                                 int(lower_n_ambiguous), _ptr(slab), slab.nbytes, _ptr(offsets), ctypes.byref(need)))
    return slab[:nbytes], offsets, lengths


def pack_q6(slab8, out=None, n_threads: int = 0):
    """6-bit transport image of a Q8 slab (moira_pack_q6): 3/4 of the bytes over PCIe.  Raises
    MoiraError(ERR_BAD_QUALITY) if a quality above 60 is present."""
    slab8 = _as(slab8, np.uint8)
    if slab8.nbytes % 16:
        slab8 = np.concatenate([slab8, np.full(16 - slab8.nbytes % 16, 0xFD, np.uint8)])
    need = slab8.nbytes // 16 * 12
    if out is None:
        out = np.empty(need, dtype=np.uint8)
    L.check(lib.moira_pack_q6(_ptr(slab8), slab8.nbytes, _ptr(out), out.nbytes, int(n_threads)))
    return out[:need]


def parse_fastq(text: bytes, fastq_offset: int = 33, lower_n_ambiguous: bool = True):
    """FASTQ bytes -> (slab, offsets, lengths, hdr_off, hdr_len, seq_off, qual_off) via moira_parse_fastq.
    Rows parsed by different host threads are separated by a little slack in the slab (offsets skip it).
    Raises MoiraError(ERR_PARSE) with the reference's error class name in the message
    (EmptySeqError / EmptyQualError / LengthMismatchError, moira.py:1178-1183)."""
    buf = np.frombuffer(text, dtype=np.uint8)
    n = ctypes.c_uint64()
    nb = ctypes.c_uint64()
    L.check(lib.moira_parse_fastq(_ptr(buf), buf.nbytes, int(fastq_offset), int(lower_n_ambiguous), None, 0,
                                  None, None, None, None, None, None, 0, ctypes.byref(n), ctypes.byref(nb)))
    nr = n.value
    slab = np.empty(max(16, nb.value), dtype=np.uint8)
    offsets = np.empty(nr, np.uint64)
    lengths = np.empty(nr, np.uint32)
    hdr_off = np.empty(nr, np.uint64)
    hdr_len = np.empty(nr, np.uint32)
    seq_off = np.empty(nr, np.uint64)
    qual_off = np.empty(nr, np.uint64)
    L.check(lib.moira_parse_fastq(_ptr(buf), buf.nbytes, int(fastq_offset), int(lower_n_ambiguous), _ptr(slab),
                                  slab.nbytes, _ptr(offsets), _ptr(lengths), _ptr(hdr_off), _ptr(hdr_len),
                                  _ptr(seq_off), _ptr(qual_off), nr, ctypes.byref(n), ctypes.byref(nb)))
    slab = slab[:max(16, nb.value)]   # the sizing call returns an upper bound, the fill call the bytes in use
    return slab, offsets, lengths, hdr_off, hdr_len, seq_off, qual_off


def parse_fasta_qual(fasta: bytes, qual: bytes, lower_n_ambiguous: bool = True):
    """FASTA + QUAL bytes -> (slab, qual_slab, offsets, lengths, hdr_off, hdr_len, seq_off) via
    moira_parse_fasta_qual.  qual_slab holds the plain qualities (Q <= 0 -> 1) at the slab's offsets."""
    fb = np.frombuffer(fasta, dtype=np.uint8)
    qb = np.frombuffer(qual, dtype=np.uint8)
    n = ctypes.c_uint64()
    nb = ctypes.c_uint64()
    L.check(lib.moira_parse_fasta_qual(_ptr(fb), fb.nbytes, _ptr(qb), qb.nbytes, int(lower_n_ambiguous), None, 0, None,
                                       None, None, None, None, None, 0, ctypes.byref(n), ctypes.byref(nb)))
    nr = n.value
    cap = max(16, nb.value)
    slab = np.empty(cap, np.uint8)
    qslab = np.empty(cap, np.uint8)
    offsets = np.empty(nr, np.uint64)
    lengths = np.empty(nr, np.uint32)
    hdr_off = np.empty(nr, np.uint64)
    hdr_len = np.empty(nr, np.uint32)
    seq_off = np.empty(nr, np.uint64)
    L.check(lib.moira_parse_fasta_qual(_ptr(fb), fb.nbytes, _ptr(qb), qb.nbytes, int(lower_n_ambiguous), _ptr(slab), cap,
                                       _ptr(qslab), _ptr(offsets), _ptr(lengths), _ptr(hdr_off), _ptr(hdr_len),
                                       _ptr(seq_off), nr, ctypes.byref(n), ctypes.byref(nb)))
    used = max(16, nb.value)
    return slab[:used], qslab[:used], offsets, lengths, hdr_off, hdr_len, seq_off


@dataclass
class CollapseResult:
    group_of_read: np.ndarray   # uint64[n]
    rep: np.ndarray             # uint64[G] representative read of every group
    size: np.ndarray            # uint64[G]
    member_start: np.ndarray    # uint64[G + 1]
    members: np.ndarray         # uint64[n] reads of group g = members[member_start[g]:member_start[g+1]] in names order
    order: np.ndarray           # uint64[G] groups by abundance (largest first, ties by first appearance)


def collapse(text, seq_off, seq_len, ee, n_threads: int = 0) -> CollapseResult:
    """moira_collapse: dereplicate identical sequences with the reference's --collapse semantics
    (moira.py:459-475, 491-504).  `text` is a bytes-like buffer holding the sequences."""
    buf = np.frombuffer(text, dtype=np.uint8) if not isinstance(text, np.ndarray) else text
    seq_off = _as(seq_off, np.uint64)
    seq_len = _as(seq_len, np.uint32)
    ee = _as(ee, np.float64)
    n = int(seq_len.shape[0])
    g_of = np.empty(n, np.uint64)
    rep = np.empty(n, np.uint64)
    size = np.empty(n, np.uint64)
    mstart = np.empty(n + 1, np.uint64)
    members = np.empty(n, np.uint64)
    order = np.empty(n, np.uint64)
    ng = ctypes.c_uint64()
    L.check(lib.moira_collapse(_ptr(buf), _ptr(seq_off), _ptr(seq_len), _ptr(ee), n, int(n_threads), _ptr(g_of),
                               ctypes.byref(ng), _ptr(rep), _ptr(size), _ptr(mstart), _ptr(members), _ptr(order)))
    G = ng.value
    return CollapseResult(g_of, rep[:G], size[:G], mstart[:G + 1], members, order[:G])


__all__ = ["collapse", "CollapseResult","Context", "FilterParams", "FilterResult", "PinnedBuffer", "MoiraError", "pack_reads",
           "pack_arrays", "pack_q6", "parse_fastq", "parse_fasta_qual", "build_lut"]
