"""Batched host API over libmoira_b200.so -- the call a host program makes instead of looping
`bernoulli.calculate_errors_PB` (reference: moira/moira.py:815-831 per read, moira.py:925-970 for
the decision).  Everything numeric happens in the CUDA library; this module only marshals buffers.
"""
from __future__ import annotations

import ctypes
import weakref
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
    cascade: int = 0                       # decisions needing 3..8 PMF entries: 0 = two-entry sweep first when the pilot says it pays, 1 = always, 2 = never
    max_length: int = 0                    # filter_device with device lengths: longest / shortest read if known (sizes the first pass)
    min_length: int = 0

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
        p.cascade = int(self.cascade)
        p.max_length, p.min_length = int(self.max_length), int(self.min_length)
        p.slab_format = L.SLAB_Q6 if self.slab_format == "q6" else L.SLAB_Q8
        p.alpha = float(self.alpha)
        return p

    @property
    def lower_n_ambiguous(self) -> bool:
        """The C core skips 'N' and 'n' (bernoullimodule.c:196); the Python calculators only 'N'
        (moira.py:1605, 1660)."""
        return self.error_calc == "poisson_binomial"


_CONSENSUS = {"best": L.CONSENSUS_BEST, "sum": L.CONSENSUS_SUM, "posterior": L.CONSENSUS_POSTERIOR}


@dataclass
class ContigParams:
    """Contig construction options, with the reference's names and defaults (moira.py:629-646)."""
    match: int = 1
    mismatch: int = -1
    gap: int = -2
    insert: int = 20
    deltaq: int = 6
    consensus_qscore: str = "best"
    qscore_cap: int = 40
    trim_overlap: bool = False

    def to_c(self) -> L.ContigParams:
        if self.consensus_qscore not in _CONSENSUS:
            raise ValueError('consensus_qscore must be "best", "sum" or "posterior".')          # moira.py:1405-1406
        p = L.ContigParams()
        p.match, p.mismatch, p.gap = int(self.match), int(self.mismatch), int(self.gap)
        p.insert, p.deltaq = int(self.insert), int(self.deltaq)
        p.consensus = _CONSENSUS[self.consensus_qscore]
        p.qscore_cap = int(self.qscore_cap)
        p.trim_overlap = 1 if self.trim_overlap else 0
        return p


@dataclass
class PairResult:
    """Contigs of a batch of read pairs (row r of contig_seq / contig_qual holds contig_len[r] entries)."""
    contig_seq: np.ndarray     # uint8[n, out_stride], ASCII
    contig_qual: np.ndarray    # uint8[n, out_stride]
    contig_len: np.ndarray     # uint32[n]
    overlap: np.ndarray        # int32[n]
    gaps: np.ndarray
    mismatches: np.ndarray
    status: np.ndarray         # uint8[n], L.PAIR_*
    filter: "FilterResult | None" = None
    _pinned: object = None

    @staticmethod
    def allocate(n: int, stride: int, with_filter: bool, pinned: bool = False) -> "PairResult":
        """Output arrays for n pairs; pinned=True puts the two big contig arrays in page-locked memory (faster D2H,
        worth it when the result object is reused across calls)."""
        if pinned:
            keep = [PinnedBuffer(max(1, n * stride)), PinnedBuffer(max(1, n * stride))]
            seq, qual = keep[0].u8[:n * stride].reshape(n, stride), keep[1].u8[:n * stride].reshape(n, stride)
        else:
            keep = None
            seq, qual = np.zeros((n, stride), np.uint8), np.zeros((n, stride), np.uint8)
        res = PairResult(seq, qual, np.zeros(n, np.uint32), np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.int32),
                         np.zeros(n, np.uint8))
        res._pinned = keep
        if with_filter:
            res.filter = FilterResult(np.zeros(n, np.float64), np.zeros(n, np.int32), np.zeros(n, np.uint8),
                                      np.zeros(L.N_COUNTERS, np.uint64))
        return res

    def contig(self, r: int):
        n = int(self.contig_len[r])
        q = self.contig_qual[r, :n].astype(np.int32)
        return self.contig_seq[r, :n].tobytes().decode("latin-1"), np.where(q > 0xFC, q - 256, q).tolist()


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

    def last_contig_ms(self):
        ms = ctypes.c_float()
        n = ctypes.c_int()
        L.check(lib.moira_ctx_last_contig_ms(self._h, ctypes.byref(ms), ctypes.byref(n)))
        return ms.value, n.value

    def fp64_peak(self, iters: int = 20000):
        ops = ctypes.c_double()
        ms = ctypes.c_double()
        L.check(lib.moira_fp64_peak(self._h, int(iters), ctypes.byref(ops), ctypes.byref(ms)))
        return ops.value, ms.value

    def link_probe(self, nbytes: int = 1 << 30, reps: int = 3):
        """(H2D GB/s, D2H GB/s) from / to pinned host memory right now (moira_link_probe)."""
        h2d, d2h = ctypes.c_double(), ctypes.c_double()
        L.check(lib.moira_link_probe(self._h, int(nbytes), int(reps), ctypes.byref(h2d), ctypes.byref(d2h)))
        return h2d.value, d2h.value

    # ---- multi-GPU: the counters' all-reduce (NCCL inside the library) -----------------------------
    def comm_init(self, unique_id: bytes, rank: int, n_ranks: int):
        """Join the communicator `unique_id` (comm_unique_id() of rank 0) as `rank` of `n_ranks`: collective."""
        if len(unique_id) != L.COMM_ID_BYTES:
            raise ValueError("unique_id must be %d bytes" % L.COMM_ID_BYTES)
        buf = (ctypes.c_uint8 * L.COMM_ID_BYTES).from_buffer_copy(unique_id)
        L.check(lib.moira_comm_init(self._h, buf, int(rank), int(n_ranks)))

    def comm_info(self):
        r, n = ctypes.c_int(), ctypes.c_int()
        L.check(lib.moira_comm_info(self._h, ctypes.byref(r), ctypes.byref(n)))
        return r.value, n.value

    def reduce_counters(self, counters: np.ndarray) -> np.ndarray:
        """Sum of the host counters over all ranks, in place (moira_reduce_counters): collective, blocking."""
        if counters.dtype != np.uint64 or counters.shape != (L.N_COUNTERS,) or not counters.flags.c_contiguous:
            raise ValueError("counters must be a contiguous uint64[%d] array" % L.N_COUNTERS)
        L.check(lib.moira_reduce_counters(self._h, _ptr(counters)))
        return counters

    def reduce_counters_device(self, d_counters: int, stream: int | None = None):
        """In-place all-reduce of a device counters array, enqueued on `stream` (moira_reduce_counters_device)."""
        L.check(lib.moira_reduce_counters_device(self._h, d_counters, stream))

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
                      d_flags: int | None, d_counters: int | None, stream: int | None = None, d_row_marks: int | None = None):
        """Device pointers (ints, e.g. torch.Tensor.data_ptr()); enqueues on `stream`, no sync.  d_row_marks: the slab's
        row marks (count_marks_device, or the slab's producer) -- the sweeps then count no N/n."""
        cp = params.to_c()
        L.check(lib.moira_filter_device(self._h, d_slab, d_offsets, d_lengths, int(stride), int(fixed_length),
                                        int(n_reads), ctypes.byref(cp), d_row_marks, d_ee, d_ns, d_flags, d_counters, stream))

    def count_marks_device(self, d_slab: int, d_offsets: int | None, d_lengths: int | None, stride: int, fixed_length: int,
                           n_reads: int, d_row_marks: int, truncate: int = 0, stream: int | None = None):
        """Row marks (Ns | has-N << 31 per row) of a device-resident slab (moira_count_marks_device); enqueues on `stream`."""
        L.check(lib.moira_count_marks_device(self._h, d_slab, d_offsets, d_lengths, int(stride), int(fixed_length), int(n_reads),
                                             int(truncate), d_row_marks, stream))

    def filter_fastq(self, text, params: FilterParams, fastq_offset: int = 33, out: FilterResult | None = None):
        """FASTQ text (bytes, or a uint8 array -- e.g. a view of pinned memory) -> FilterResult in one streaming C call
        (moira_filter_fastq); also returns the read lengths.  With `out` given, its size is the read capacity and the
        counting pass over the text is skipped."""
        buf = text if isinstance(text, np.ndarray) else np.frombuffer(text, dtype=np.uint8)
        n = ctypes.c_uint64()
        if out is None:
            L.check(lib.moira_fastq_count_reads(_ptr(buf), buf.nbytes, ctypes.byref(n)))
            nr = n.value
            out = FilterResult(np.empty(nr, np.float64), np.empty(nr, np.int32), np.empty(nr, np.uint8),
                               np.zeros(L.N_COUNTERS, np.uint64))
        else:
            nr = int(out.ee.shape[0])
        lengths = np.empty(nr, np.uint32)
        cp = params.to_c()
        L.check(lib.moira_filter_fastq(self._h, _ptr(buf), buf.nbytes, int(fastq_offset), int(params.lower_n_ambiguous),
                                       ctypes.byref(cp), nr, _ptr(out.ee), _ptr(out.ns), _ptr(out.flags), _ptr(lengths),
                                       _ptr(out.counters), ctypes.byref(n)))
        return out, lengths[:n.value]

    def filter_fastq_ex(self, text, params: FilterParams, fastq_offset: int = 33, max_reads: int | None = None,
                        offsets: bool = True, labels: bool = False) -> "FastqResult":
        """moira_filter_fastq_ex: decisions plus what the host needs to go on to the output files without parsing again
        -- per read the position of its sequence and quality line in `text`, and (labels=True) the device-side
        dereplication of the (truncated) sequences."""
        buf = text if isinstance(text, np.ndarray) else np.frombuffer(text, dtype=np.uint8)
        if max_reads is None:
            n = ctypes.c_uint64()
            L.check(lib.moira_fastq_count_reads(_ptr(buf), buf.nbytes, ctypes.byref(n)))
            max_reads = n.value
        nr = int(max_reads)
        res = FilterResult(np.empty(nr, np.float64), np.empty(nr, np.int32), np.empty(nr, np.uint8), np.zeros(L.N_COUNTERS, np.uint64))
        lengths = np.empty(nr, np.uint32)
        soff = np.empty(nr, np.uint64) if offsets else None
        qoff = np.empty(nr, np.uint64) if offsets else None
        lab = np.empty(nr, np.uint32) if labels else None
        got = ctypes.c_uint64()
        cp = params.to_c()
        L.check(lib.moira_filter_fastq_ex(self._h, _ptr(buf), buf.nbytes, int(fastq_offset), int(params.lower_n_ambiguous),
                                          ctypes.byref(cp), nr, _ptr(res.ee), _ptr(res.ns), _ptr(res.flags), _ptr(lengths),
                                          _ptr(soff), _ptr(qoff), _ptr(lab), _ptr(res.counters), ctypes.byref(got)))
        m = got.value
        res = FilterResult(res.ee[:m], res.ns[:m], res.flags[:m], res.counters)
        return FastqResult(res, lengths[:m], None if soff is None else soff[:m], None if qoff is None else qoff[:m],
                           None if lab is None else lab[:m])

    def collapse_device(self, d_seq: int, d_offsets: int | None, d_lengths: int | None, stride: int, fixed_length: int,
                        n_reads: int, d_labels: int, truncate: int = 0, stream: int | None = None):
        """Device-resident sequences -> labels (moira_collapse_device); enqueues on `stream`."""
        L.check(lib.moira_collapse_device(self._h, d_seq, d_offsets, d_lengths, int(stride), int(fixed_length), int(n_reads),
                                          int(truncate), d_labels, stream))

    def collapse_labels(self, labels, ee) -> "CollapseResult":
        """moira_collapse_labels_device: groups, representatives, names order and abundance order from labels and ee, computed
        on this context's GPU (host arrays in and out); identical to the module-level collapse_labels (host)."""
        labels = _as(labels, np.uint32)
        ee = _as(ee, np.float64)
        n = int(labels.shape[0])
        g_of, rep, size = np.empty(n, np.uint64), np.empty(n, np.uint64), np.empty(n, np.uint64)
        mstart, members, order = np.empty(n + 1, np.uint64), np.empty(n, np.uint64), np.empty(n, np.uint64)
        ng = ctypes.c_uint64()
        L.check(lib.moira_collapse_labels_device(self._h, _ptr(labels), _ptr(ee), n, _ptr(g_of), ctypes.byref(ng), _ptr(rep),
                                                 _ptr(size), _ptr(mstart), _ptr(members), _ptr(order)))
        G = ng.value
        return CollapseResult(g_of, rep[:G], size[:G], mstart[:G + 1], members, order[:G])

    def collapse_addr(self, seq_addr, seq_len, truncate: int = 0) -> np.ndarray:
        """moira_collapse_addr: dereplication labels (uint32[n]) of sequences given by absolute host addresses, made on the GPU."""
        seq_addr, seq_len = _as(seq_addr, np.uint64), _as(seq_len, np.uint32)
        n = int(seq_addr.shape[0])
        labels = np.empty(n, np.uint32)
        L.check(lib.moira_collapse_addr(self._h, _ptr(seq_addr), _ptr(seq_len), n, int(truncate or 0), _ptr(labels)))
        return labels

    def collapse_groups(self, labels, ee, n: int | None = None) -> "CollapseResult":
        """moira_collapse_groups: as collapse_labels, without the widening copy -- the result arrays are uint32 views of the
        context's pinned memory (valid until the next collapse call on this context).  labels / ee: numpy arrays, or device
        pointers (ints) together with n."""
        on_device = isinstance(labels, int)
        if on_device:
            lp, ep, n = ctypes.c_void_p(labels), ctypes.c_void_p(ee), int(n)
        else:
            labels, ee = _as(labels, np.uint32), _as(ee, np.float64)
            lp, ep, n = _ptr(labels), _ptr(ee), int(labels.shape[0])
        ng = ctypes.c_uint64()
        ptrs = [ctypes.c_void_p() for _ in range(6)]
        L.check(lib.moira_collapse_groups(self._h, lp, ep, 1 if on_device else 0, n, ctypes.byref(ng), *[ctypes.byref(p) for p in ptrs]))
        G = ng.value
        counts = [n, G, G, G + 1, n, G]

        def view(p, c):
            if not c or not p.value:
                return np.empty(0, np.uint32)
            return np.ctypeslib.as_array((ctypes.c_uint32 * c).from_address(p.value))
        return CollapseResult(*[view(p, c) for p, c in zip(ptrs, counts)])

    def filter_pairs(self, fwd_seq, fwd_qual, fwd_off, fwd_len, rev_seq, rev_qual, rev_off, rev_len,
                     contig_params: ContigParams, filter_params: FilterParams | None = None,
                     lower_n_ambiguous: bool = True, fwd_qual_off=None, rev_qual_off=None, qual_base: int = 0,
                     out: "PairResult | None" = None) -> PairResult:
        """Read pairs -> contigs (and, with filter_params, the filter on them) in one C call (moira_filter_pairs).
        *_seq: uint8 ASCII bases at [off[r], off[r] + len[r]); *_qual: uint8 qualities (+ qual_base) at
        [qual_off[r], ...) (qual_off None: same offsets).  Passing the FASTQ text as both arrays with the parser's
        seq_off / qual_off and qual_base = the FASTQ offset avoids any repacking."""
        fwd_seq, fwd_qual = _as(fwd_seq, np.uint8), _as(fwd_qual, np.uint8)
        rev_seq, rev_qual = _as(rev_seq, np.uint8), _as(rev_qual, np.uint8)
        fwd_off, rev_off = _as(fwd_off, np.uint64), _as(rev_off, np.uint64)
        fwd_len, rev_len = _as(fwd_len, np.uint32), _as(rev_len, np.uint32)
        fwd_qual_off = None if fwd_qual_off is None else _as(fwd_qual_off, np.uint64)
        rev_qual_off = None if rev_qual_off is None else _as(rev_qual_off, np.uint64)
        n = int(fwd_len.shape[0])
        if not (rev_len.shape[0] == fwd_off.shape[0] == rev_off.shape[0] == n):
            raise ValueError("offset / length arrays of the two files differ in size")
        mx = (int(fwd_len.max()) if n else 0) + (min(int(rev_len.max()), 1024) if n else 0)
        stride = max(16, (mx + 15) // 16 * 16)
        if out is not None:
            if out.contig_seq.shape != (n, stride) or (filter_params is not None) != (out.filter is not None):
                raise ValueError("`out` was allocated for another batch shape")
            res = out
        else:
            res = PairResult.allocate(n, stride, filter_params is not None)
        cp = contig_params.to_c()
        fp = filter_params.to_c() if filter_params is not None else None
        fr = res.filter
        L.check(lib.moira_filter_pairs(
            self._h, _ptr(fwd_seq), fwd_seq.nbytes, _ptr(fwd_qual), fwd_qual.nbytes, _ptr(fwd_off), _ptr(fwd_qual_off),
            _ptr(fwd_len), _ptr(rev_seq), rev_seq.nbytes, _ptr(rev_qual), rev_qual.nbytes, _ptr(rev_off), _ptr(rev_qual_off),
            _ptr(rev_len), int(qual_base), n, ctypes.byref(cp), int(lower_n_ambiguous),
            ctypes.byref(fp) if fp is not None else None, stride, _ptr(res.contig_seq), _ptr(res.contig_qual),
            _ptr(res.contig_len), _ptr(res.overlap), _ptr(res.gaps), _ptr(res.mismatches), _ptr(res.status),
            _ptr(fr.ee) if fr else None, _ptr(fr.ns) if fr else None, _ptr(fr.flags) if fr else None,
            _ptr(fr.counters) if fr else None))
        return res

    def nw_align(self, seq_1: str, seq_2: str, match: int, mismatch: int, gap: int):
        """One alignment through the CUDA path (moira_nw_align) -> (seq_1_aligned, seq_2_aligned, score)."""
        cap = len(seq_1) + len(seq_2) + 1
        a1, a2 = ctypes.create_string_buffer(cap), ctypes.create_string_buffer(cap)
        n, score = ctypes.c_uint64(), ctypes.c_int64()
        L.check(lib.moira_nw_align(self._h, seq_1.encode("latin-1"), seq_2.encode("latin-1"), int(match), int(mismatch), int(gap),
                                   a1, a2, ctypes.byref(n), ctypes.byref(score)))
        return a1.raw[:n.value].decode("latin-1"), a2.raw[:n.value].decode("latin-1"), score.value

    def make_contig(self, forward_aligned: str, forward_quals, reverse_aligned: str, reverse_quals, params: ContigParams):
        """One consensus through the CUDA path (moira_make_contig) -> (contig, quals, overlap_length, gaps, mismatches)."""
        fq = np.ascontiguousarray(forward_quals, dtype=np.int32)
        rq = np.ascontiguousarray(reverse_quals, dtype=np.int32)
        cap = len(forward_aligned) + 1
        contig = ctypes.create_string_buffer(cap)
        cq = np.zeros(cap, np.int32)
        n = ctypes.c_uint64()
        ov, gp, mm = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
        cp = params.to_c()
        L.check(lib.moira_make_contig(self._h, forward_aligned.encode("latin-1"), _ptr(fq) if fq.size else None, int(fq.size),
                                      reverse_aligned.encode("latin-1"), _ptr(rq) if rq.size else None, int(rq.size),
                                      ctypes.byref(cp), contig, _ptr(cq), ctypes.byref(n), ctypes.byref(ov), ctypes.byref(gp),
                                      ctypes.byref(mm)))
        return contig.raw[:n.value].decode("latin-1"), cq[:n.value].tolist(), ov.value, gp.value, mm.value

    def calculate_errors_PB(self, contig: str, contig_quals, alpha: float):
        """One read through the CUDA path (moira_calculate_errors_PB)."""
        q = np.ascontiguousarray(contig_quals, dtype=np.int32)
        ee = ctypes.c_double()
        ns = ctypes.c_int32()
        rc = lib.moira_calculate_errors_PB(self._h, contig.encode("latin-1"), _ptr(q) if q.size else None,
                                           int(q.shape[0]), float(alpha), ctypes.byref(ee), ctypes.byref(ns))
        L.check(rc)
        return ee.value, ns.value


def comm_unique_id() -> bytes:
    """A fresh communicator id (moira_comm_unique_id): rank 0 makes it, every rank passes it to Context.comm_init."""
    buf = (ctypes.c_uint8 * L.COMM_ID_BYTES)()
    L.check(lib.moira_comm_unique_id(buf))
    return bytes(buf)


def comm_init_all(contexts):
    """One process, several GPUs: a communicator over `contexts` (distinct devices), rank i = contexts[i]."""
    arr = (ctypes.c_void_p * len(contexts))(*[c._h for c in contexts])
    L.check(lib.moira_comm_init_all(arr, len(contexts)))


def reduce_counters_all(contexts, counters):
    """One process, several GPUs: every counters[i] (uint64[N_COUNTERS], context i) becomes the sum over all contexts
    (one NCCL group call, moira_reduce_counters_all)."""
    arr = (ctypes.c_void_p * len(contexts))(*[c._h for c in contexts])
    ptrs = (ctypes.c_void_p * len(contexts))(*[c_.ctypes.data for c_ in counters])
    L.check(lib.moira_reduce_counters_all(arr, len(contexts), ptrs))
    return counters


def pack_sequences(seqs, quals_list):
    """Python reads -> (bases uint8, qualities uint8, offsets uint64, lengths uint32), back to back: the layout
    Context.filter_pairs takes for either file of a pair."""
    lengths = np.array([len(s) for s in seqs], dtype=np.uint32)
    offsets = np.zeros(len(seqs), dtype=np.uint64)
    if len(seqs):
        offsets[1:] = np.cumsum(lengths.astype(np.uint64))[:-1]
    bases = np.frombuffer("".join(seqs).encode("latin-1"), dtype=np.uint8).copy() if len(seqs) else np.zeros(0, np.uint8)
    flat = [q for ql in quals_list for q in ql]
    if any(q < 0 or q > 0xFC for q in flat):
        raise MoiraError(L.ERR_BAD_QUALITY, "quality outside 0..252")
    quals = np.asarray(flat, dtype=np.uint8)
    if bases.size == 0:
        bases, quals = np.zeros(1, np.uint8), np.zeros(1, np.uint8)
    return bases, quals, offsets, lengths


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


def index_fastq(text):
    """FASTQ bytes -> (lengths, hdr_off, hdr_len, seq_off, qual_off) via moira_index_fastq: the record table without a slab,
    for flows that hand the text itself to the device (read pairs).  Errors as parse_fastq."""
    buf = np.frombuffer(text, dtype=np.uint8)
    n = ctypes.c_uint64()
    L.check(lib.moira_index_fastq(_ptr(buf), buf.nbytes, None, None, None, None, None, 0, ctypes.byref(n)))
    nr = n.value
    lengths, hdr_len = np.empty(nr, np.uint32), np.empty(nr, np.uint32)
    hdr_off, seq_off, qual_off = np.empty(nr, np.uint64), np.empty(nr, np.uint64), np.empty(nr, np.uint64)
    if nr:
        L.check(lib.moira_index_fastq(_ptr(buf), buf.nbytes, _ptr(lengths), _ptr(hdr_off), _ptr(hdr_len), _ptr(seq_off), _ptr(qual_off),
                                      nr, ctypes.byref(n)))
    return lengths, hdr_off, hdr_len, seq_off, qual_off


def parse_fasta_qual(fasta: bytes, qual: bytes, lower_n_ambiguous: bool = True):
    """FASTA + QUAL bytes -> (slab, qual_slab, offsets, lengths, hdr_off, hdr_len, seq_off) via
    moira_parse_fasta_qual.  qual_slab holds the plain qualities (negative values as 0) at the slab's offsets."""
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
    buf = None if text is None else (np.frombuffer(text, dtype=np.uint8) if not isinstance(text, np.ndarray) else text)   # None: absolute addresses
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


def collapse_labels(labels, ee) -> CollapseResult:
    """moira_collapse_labels: the reference's groups, representatives, names order and abundance order from labels
    (equal label <=> equal sequence) and ee."""
    labels = _as(labels, np.uint32)
    ee = _as(ee, np.float64)
    n = int(labels.shape[0])
    g_of, rep, size = np.empty(n, np.uint64), np.empty(n, np.uint64), np.empty(n, np.uint64)
    mstart, members, order = np.empty(n + 1, np.uint64), np.empty(n, np.uint64), np.empty(n, np.uint64)
    ng = ctypes.c_uint64()
    L.check(lib.moira_collapse_labels(_ptr(labels), _ptr(ee), n, _ptr(g_of), ctypes.byref(ng), _ptr(rep), _ptr(size), _ptr(mstart),
                                      _ptr(members), _ptr(order)))
    G = ng.value
    return CollapseResult(g_of, rep[:G], size[:G], mstart[:G + 1], members, order[:G])


def _gz_owned_array(ptr: int, nbytes: int):
    """uint8 array over a buffer of moira_gz_inflate; the buffer is released when the last view of it is gone."""
    if not nbytes:
        lib.moira_gz_free(ptr)
        return np.zeros(0, np.uint8)
    carr = (ctypes.c_uint8 * nbytes).from_address(ptr)
    weakref.finalize(carr, lib.moira_gz_free, ctypes.c_void_p(ptr))
    return np.frombuffer(carr, dtype=np.uint8)


def gz_scan(gz):
    """(members, inflated bytes) of a BGZF file -- gzip members that carry their own size: read without inflating --, (0, 0)
    for any other file."""
    a = np.frombuffer(gz, dtype=np.uint8) if not isinstance(gz, np.ndarray) else gz
    n, total = ctypes.c_uint64(), ctypes.c_uint64()
    L.check(lib.moira_gz_scan(_ptr(a), a.size, ctypes.byref(n), ctypes.byref(total)))
    return n.value, total.value


def gz_inflate(gz, n_threads: int = 0):
    """gzip.GzipFile(...).read() on the host threads (moira.py:1065-1068): BGZF / blocked gzip members in parallel, any other
    gzip file on one thread.  Returns a uint8 array over memory the library owns (released with its last view).
    MoiraError(ERR_PARSE) for files that are not gzip, truncated or corrupt."""
    a = np.frombuffer(gz, dtype=np.uint8) if not isinstance(gz, np.ndarray) else gz
    out, nb = ctypes.c_void_p(), ctypes.c_uint64()
    L.check(lib.moira_gz_inflate(_ptr(a), a.size, int(n_threads), ctypes.byref(out), ctypes.byref(nb)))
    return _gz_owned_array(out.value, nb.value)


def gz_deflate(data, fd: int, offset: int = 0, level: int = 6, n_threads: int = 0, eof: bool = False) -> int:
    """`data` -> BGZF members at byte `offset` of the open descriptor fd (parallel compression); returns the bytes written."""
    a = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    n = ctypes.c_uint64()
    L.check(lib.moira_gz_deflate(_ptr(a), a.size, int(level), int(n_threads), int(fd), int(offset), ctypes.byref(n)))
    total = n.value
    if eof:
        L.check(lib.moira_gz_eof(int(fd), int(offset) + total, ctypes.byref(n)))
        total += n.value
    return total


def fastq_headers(text, seq_off, n_threads: int = 0):
    """(hdr_off uint64[n], hdr_len uint32[n]): header tokens of FASTQ records from their sequence-line positions."""
    buf = text if isinstance(text, np.ndarray) else np.frombuffer(text, dtype=np.uint8)
    seq_off = _as(seq_off, np.uint64)
    n = int(seq_off.shape[0])
    ho, hl = np.empty(n, np.uint64), np.empty(n, np.uint32)
    L.check(lib.moira_fastq_headers(_ptr(buf), buf.nbytes, _ptr(seq_off), n, int(n_threads), _ptr(ho), _ptr(hl)))
    return ho, hl


def line_offsets(text, line_numbers):
    """(offsets uint64[k], n_lines): byte offsets just behind the given (ascending) numbers of newlines of `text`, and its
    newline count (moira_line_offsets)."""
    buf = text if isinstance(text, np.ndarray) else np.frombuffer(text, dtype=np.uint8)
    q = _as(line_numbers, np.uint64)
    out = np.empty(q.shape[0], np.uint64)
    n = ctypes.c_uint64()
    L.check(lib.moira_line_offsets(_ptr(buf), buf.nbytes, _ptr(q), int(q.shape[0]), _ptr(out), ctypes.byref(n)))
    return out, n.value


def fastq_split(text, n_parts: int):
    """Record-aligned cut points (uint64[n_parts + 1]) of a FASTQ text, one shard per GPU (moira_fastq_split)."""
    buf = text if isinstance(text, np.ndarray) else np.frombuffer(text, dtype=np.uint8)
    cuts = np.zeros(n_parts + 1, np.uint64)
    L.check(lib.moira_fastq_split(_ptr(buf), buf.nbytes, int(n_parts), _ptr(cuts)))
    return cuts


@dataclass
class FastqResult:
    filter: FilterResult
    lengths: np.ndarray            # uint32[n]
    seq_off: np.ndarray | None     # uint64[n] position of the sequence line in the text
    qual_off: np.ndarray | None
    labels: np.ndarray | None      # uint32[n] read index of one read with the same (truncated) sequence


@dataclass
class RecordView:
    """Where the reads' text lives (struct moira_records): base buffers (uint8 arrays) + per-read byte ranges."""
    hdr_base: np.ndarray
    hdr_off: np.ndarray
    hdr_len: np.ndarray
    seq_base: np.ndarray
    seq_off: np.ndarray
    qual_base: np.ndarray
    qual_off: np.ndarray
    length: np.ndarray             # bases to write (after --truncate)
    qual_sub: int = 0


class Blocks:
    """Formatted output (moira_format_records): memory blocks owned by the library until close()."""

    def __init__(self, handle):
        self._h = handle
        n = ctypes.c_int()
        L.check(lib.moira_blocks_parts(self._h, ctypes.byref(n)))
        self.n_parts = n.value

    def block(self, part: int, which: int):
        ptr, ln = ctypes.c_void_p(), ctypes.c_uint64()
        L.check(lib.moira_blocks_get(self._h, part, which, ctypes.byref(ptr), ctypes.byref(ln)))
        if not ln.value:
            return b""
        return (ctypes.c_char * ln.value).from_address(ptr.value)

    def write(self, which: int, fh):
        """All parts of one block kind, in order, to a binary file object."""
        for p in range(self.n_parts):
            b = self.block(p, which)
            if len(b):
                fh.write(b)

    def pwrite(self, which: int, fd: int, offset: int, n_threads: int = 0) -> int:
        """All parts of one block kind to file descriptor fd at byte `offset` (native, parallel); returns the bytes written."""
        n = ctypes.c_uint64()
        L.check(lib.moira_blocks_write(self._h, int(which), int(fd), int(offset), int(n_threads), ctypes.byref(n)))
        return n.value

    def pwrite_gz(self, which: int, fd: int, offset: int, level: int = 6, n_threads: int = 0) -> int:
        """The same through gzip: BGZF members compressed on all host threads (moira_blocks_write_gz); returns the bytes written."""
        n = ctypes.c_uint64()
        L.check(lib.moira_blocks_write_gz(self._h, int(which), int(fd), int(offset), int(level), int(n_threads), ctypes.byref(n)))
        return n.value

    def recycle(self):
        """Hand the memory back to the library for the next format_records call (steady streams of batches)."""
        if self._h:
            lib.moira_blocks_recycle(self._h)
            self._h = None

    def close(self):
        if self._h:
            lib.moira_blocks_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def format_records(rec: RecordView, sel, ee, accept, reason, *, fastq: bool, fastq_offset: int = 33, usearch: bool = False,
                   names: bool = False, relabel: str | None = None, first_index: int = 0, notes=None, sel_group=None,
                   member_start=None, members=None, stats=None, n_threads: int = 0) -> Blocks:
    """moira_format_records: the records `sel` (read indices in output order, None = all) as write_results prints them
    (moira.py:842-970).  notes: {reason code: bytes appended to a rejected record's header}.  stats: (overlap, gaps,
    mismatches) int32 arrays for the contigs report."""
    keep = []

    def arr(a, dt):
        if a is None:
            return None
        a = np.ascontiguousarray(a, dtype=dt)
        keep.append(a)
        return a

    r = L.Records()
    bases = [arr(rec.hdr_base, np.uint8), arr(rec.seq_base, np.uint8), arr(rec.qual_base, np.uint8)]
    r.hdr_base, r.seq_base, r.qual_base = [_ptr(b) for b in bases]      # None: the offsets are absolute addresses
    r.hdr_off, r.hdr_len = _ptr(arr(rec.hdr_off, np.uint64)), _ptr(arr(rec.hdr_len, np.uint32))
    r.seq_off, r.qual_off = _ptr(arr(rec.seq_off, np.uint64)), _ptr(arr(rec.qual_off, np.uint64))
    r.len = _ptr(arr(rec.length, np.uint32))
    r.qual_sub = int(rec.qual_sub)
    o = L.WriteOpts()
    o.fastq, o.fastq_offset, o.usearch, o.names = int(bool(fastq)), int(fastq_offset), int(bool(usearch)), int(bool(names))
    o.relabel = relabel.encode("latin-1") if relabel else None
    o.first_index = int(first_index)
    for code, text in (notes or {}).items():
        o.notes[int(code)] = text if isinstance(text, bytes) else text.encode("latin-1")
    sel = arr(sel, np.uint64)
    n_sel = int(sel.shape[0]) if sel is not None else int(np.asarray(rec.length).shape[0])
    ov, gp, mm = (None, None, None) if stats is None else [arr(x, np.int32) for x in stats]
    out = ctypes.c_void_p()
    L.check(lib.moira_format_records(ctypes.byref(r), ctypes.byref(o), _ptr(sel), n_sel, _ptr(arr(ee, np.float64)),
                                     _ptr(arr(accept, np.uint8)), _ptr(arr(reason, np.uint8)), _ptr(arr(sel_group, np.uint64)),
                                     _ptr(arr(member_start, np.uint64)), _ptr(arr(members, np.uint64)), _ptr(ov), _ptr(gp), _ptr(mm),
                                     int(n_threads), ctypes.byref(out)))
    return Blocks(out)


__all__ = ["collapse_labels", "fastq_headers", "fastq_split", "FastqResult", "RecordView", "Blocks", "format_records", "collapse", "CollapseResult","Context", "FilterParams", "FilterResult", "PinnedBuffer", "MoiraError", "pack_reads",
           "pack_arrays", "pack_q6", "parse_fastq", "parse_fasta_qual", "build_lut", "ContigParams", "PairResult", "pack_sequences",
           "comm_unique_id", "comm_init_all", "reduce_counters_all"]
