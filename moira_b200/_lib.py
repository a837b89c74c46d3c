"""ctypes binding of libmoira_b200.so (include/moira_b200.h).

There is deliberately no fallback: if the shared library is missing this module raises at import
time, and if no CUDA device is usable every compute entry point raises MoiraError(MOIRA_ERR_CUDA).
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MOIRA_B200_LIB") or os.path.join(_HERE, "libmoira_b200.so")   # env override: tuning experiments only

# ---- constants mirrored from include/moira_b200.h -------------------------------------------
ABI_VERSION = 6
OK = 0
ERR_BAD_ALPHA, ERR_LENGTH_MISMATCH, ERR_BAD_QUALITY, ERR_CUDA = -1, -2, -3, -4
ERR_BAD_ARG, ERR_NOMEM, ERR_UNRESOLVED, ERR_PARSE, ERR_NCCL = -5, -6, -7, -8, -9
COMM_ID_BYTES = 128
MODE_PB, MODE_POISSON, MODE_EXPECTED_ERROR = 0, 1, 2
THR_UNCERT, THR_MAXERRORS = 0, 1
AMBIGS_TREAT_AS_ERRORS, AMBIGS_IGNORE, AMBIGS_DISALLOW = 0, 1, 2
EE_RAW, EE_FINAL = 0, 1
SLAB_Q8, SLAB_Q6 = 0, 1
FLAG_ACCEPT, FLAG_REASON_MASK, FLAG_LOWER_BOUND = 0x01, 0x0E, 0x10
FLAG_HAS_N, FLAG_NUMERIC, FLAG_NEAR_CUTOFF = 0x20, 0x40, 0x80
REASON_NONE, REASON_ERRORS, REASON_LENGTH, REASON_AMBIGS = 0, 1, 2, 3
CNT_READS, CNT_ACCEPTED, CNT_BAD_ERRORS, CNT_BAD_LENGTH, CNT_BAD_AMBIGS = 0, 1, 2, 3, 4
CNT_NEAR_CUTOFF, CNT_LOWER_BOUND, CNT_NUMERIC, CNT_HIST, N_HIST, N_COUNTERS = 5, 6, 7, 16, 64, 80
CNT_ESCALATED = 8
CNT_FP64_OPS = 9
CNT_CLASSIFIED = 10
MAX_INFLIGHT = 4
CONSENSUS_BEST, CONSENSUS_SUM, CONSENSUS_POSTERIOR = 0, 1, 2
PAIR_OK, PAIR_EMPTY, PAIR_BAD_BASE, PAIR_BAD_QUALITY, PAIR_TOO_LONG = 0, 1, 2, 3, 4

EXPORTS = [
    "moira_abi_version", "moira_last_error", "moira_params_default", "moira_device_count", "moira_ctx_create",
    "moira_ctx_destroy", "moira_ctx_sm_count", "moira_build_lut", "moira_ctx_get_lut",
    "moira_host_alloc", "moira_host_free", "moira_filter_device", "moira_count_marks_device", "moira_filter_batch",
    "moira_submit", "moira_wait", "moira_calculate_errors_PB", "moira_pack_reads", "moira_pack_q6",
    "moira_parse_fastq", "moira_parse_fasta_qual", "moira_fastq_count_reads", "moira_index_fastq", "moira_filter_fastq", "moira_collapse", "moira_set_host_threads", "moira_fp64_peak", "moira_ctx_launch_count", "moira_ctx_set_timing",
    "moira_filter_fastq_ex", "moira_collapse_device", "moira_collapse_labels", "moira_collapse_labels_device", "moira_collapse_groups", "moira_collapse_addr", "moira_format_records", "moira_blocks_parts",
    "moira_blocks_get", "moira_blocks_write", "moira_blocks_recycle", "moira_blocks_free", "moira_fastq_headers", "moira_fastq_split", "moira_line_offsets",
    "moira_comm_unique_id", "moira_comm_init", "moira_comm_init_all", "moira_comm_info", "moira_reduce_counters_device",
    "moira_reduce_counters", "moira_reduce_counters_all", "moira_link_probe",
    "moira_gz_scan", "moira_gz_inflate", "moira_gz_free", "moira_gz_deflate", "moira_gz_eof", "moira_blocks_write_gz",
    "moira_ctx_last_kernel_ms", "moira_ctx_last_contig_ms", "moira_contig_params_default", "moira_filter_pairs", "moira_nw_align", "moira_make_contig",
]


class Params(ctypes.Structure):
    """struct moira_params."""
    _fields_ = [
        ("mode", ctypes.c_int32), ("thr_kind", ctypes.c_int32), ("ambigs", ctypes.c_int32),
        ("round_flag", ctypes.c_int32), ("truncate", ctypes.c_uint32), ("exact_ee", ctypes.c_int32),
        ("ee_output", ctypes.c_int32), ("length_sort", ctypes.c_int32),
        ("slab_format", ctypes.c_int32), ("cascade", ctypes.c_int32),
        ("max_length", ctypes.c_uint32), ("min_length", ctypes.c_uint32),
        ("alpha", ctypes.c_double), ("thr", ctypes.c_double),
    ]


class ContigParams(ctypes.Structure):
    """struct moira_contig_params."""
    _fields_ = [(n, ctypes.c_int32) for n in ("match", "mismatch", "gap", "insert", "deltaq", "consensus", "qscore_cap",
                                              "trim_overlap")]


class Records(ctypes.Structure):
    """struct moira_records."""
    _fields_ = [("hdr_base", ctypes.c_void_p), ("hdr_off", ctypes.c_void_p), ("hdr_len", ctypes.c_void_p),
                ("seq_base", ctypes.c_void_p), ("seq_off", ctypes.c_void_p), ("qual_base", ctypes.c_void_p),
                ("qual_off", ctypes.c_void_p), ("len", ctypes.c_void_p), ("qual_sub", ctypes.c_int32)]


class WriteOpts(ctypes.Structure):
    """struct moira_write_opts."""
    _fields_ = [("fastq", ctypes.c_int32), ("fastq_offset", ctypes.c_int32), ("usearch", ctypes.c_int32), ("names", ctypes.c_int32),
                ("relabel", ctypes.c_char_p), ("first_index", ctypes.c_uint64), ("notes", ctypes.c_char_p * 8)]


BLOCK_GOOD, BLOCK_GOOD_QUAL, BLOCK_GOOD_NAMES, BLOCK_BAD, BLOCK_BAD_QUAL, BLOCK_BAD_NAMES, BLOCK_REPORT, BLOCK_N = range(8)


class MoiraError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("moira_b200 error %d: %s" % (code, message))
        self.code = code
        self.message = message


if not os.path.exists(LIB_PATH):
    raise ImportError(
        "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` or "
        "`make -C moira_b200/csrc` (there is no CPU fallback)" % LIB_PATH)

lib = ctypes.CDLL(LIB_PATH)

_vp, _i, _u64, _u32, _dbl = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_double
_pp = ctypes.POINTER(Params)

lib.moira_abi_version.restype = _i
lib.moira_last_error.restype = ctypes.c_char_p
lib.moira_params_default.restype = None
lib.moira_params_default.argtypes = [_pp]
lib.moira_device_count.argtypes = [ctypes.POINTER(_i)]
lib.moira_ctx_create.argtypes = [_i, ctypes.POINTER(_vp)]
lib.moira_ctx_destroy.argtypes = [_vp]
lib.moira_ctx_sm_count.argtypes = [_vp, ctypes.POINTER(_i)]
lib.moira_build_lut.argtypes = [_vp, _vp, _vp, ctypes.POINTER(_i)]
lib.moira_ctx_get_lut.argtypes = [_vp, _vp, _vp, _vp]
lib.moira_host_alloc.argtypes = [ctypes.POINTER(_vp), ctypes.c_size_t]
lib.moira_host_free.argtypes = [_vp]
lib.moira_filter_device.argtypes = [_vp, _vp, _vp, _vp, _u64, _u32, _u64, _pp, _vp, _vp, _vp, _vp, _vp, _vp]
lib.moira_count_marks_device.argtypes = [_vp, _vp, _vp, _vp, _u64, _u32, _u64, _u32, _vp, _vp]
lib.moira_filter_batch.argtypes = [_vp, _vp, _u64, _vp, _vp, _u64, _pp, _vp, _vp, _vp, _vp]
lib.moira_submit.argtypes = [_vp, _vp, _u64, _vp, _vp, _u64, _pp, _vp, _vp, _vp, _vp, ctypes.POINTER(_i)]
lib.moira_wait.argtypes = [_vp, _i]
lib.moira_calculate_errors_PB.argtypes = [_vp, ctypes.c_char_p, _vp, _u64, _dbl, ctypes.POINTER(_dbl),
                                          ctypes.POINTER(ctypes.c_int32)]
lib.moira_pack_reads.argtypes = [_vp, _vp, _vp, _vp, _u64, _i, _vp, _u64, _vp, ctypes.POINTER(_u64)]
lib.moira_pack_q6.argtypes = [_vp, _u64, _vp, _u64, _i]
lib.moira_parse_fastq.argtypes = [_vp, _u64, _i, _i, _vp, _u64, _vp, _vp, _vp, _vp, _vp, _vp, _u64,
                                  ctypes.POINTER(_u64), ctypes.POINTER(_u64)]
lib.moira_parse_fasta_qual.argtypes = [_vp, _u64, _vp, _u64, _i, _vp, _u64, _vp, _vp, _vp, _vp, _vp, _vp, _u64,
                                       ctypes.POINTER(_u64), ctypes.POINTER(_u64)]
lib.moira_fastq_count_reads.argtypes = [_vp, _u64, ctypes.POINTER(_u64)]
lib.moira_index_fastq.argtypes = [_vp, _u64, _vp, _vp, _vp, _vp, _vp, _u64, ctypes.POINTER(_u64)]
lib.moira_filter_fastq.argtypes = [_vp, _vp, _u64, _i, _i, _pp, _u64, _vp, _vp, _vp, _vp, _vp, ctypes.POINTER(_u64)]
lib.moira_filter_fastq_ex.argtypes = [_vp, _vp, _u64, _i, _i, _pp, _u64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, ctypes.POINTER(_u64)]
lib.moira_collapse_device.argtypes = [_vp, _vp, _vp, _vp, _u64, _u32, _u64, _u32, _vp, _vp]
lib.moira_collapse_labels.argtypes = [_vp, _vp, _u64, _vp, ctypes.POINTER(_u64), _vp, _vp, _vp, _vp, _vp]
lib.moira_collapse_addr.argtypes = [_vp, _vp, _vp, _u64, _u32, _vp]
lib.moira_collapse_groups.argtypes = [_vp, _vp, _vp, _i, _u64, ctypes.POINTER(_u64)] + [ctypes.POINTER(_vp)] * 6
lib.moira_collapse_labels_device.argtypes = [_vp, _vp, _vp, _u64, _vp, ctypes.POINTER(_u64), _vp, _vp, _vp, _vp, _vp]
lib.moira_format_records.argtypes = [ctypes.POINTER(Records), ctypes.POINTER(WriteOpts), _vp, _u64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                     _vp, _i, ctypes.POINTER(_vp)]
lib.moira_blocks_parts.argtypes = [_vp, ctypes.POINTER(_i)]
lib.moira_blocks_get.argtypes = [_vp, _i, _i, ctypes.POINTER(_vp), ctypes.POINTER(_u64)]
lib.moira_line_offsets.argtypes = [_vp, _u64, _vp, _u64, _vp, ctypes.POINTER(_u64)]
lib.moira_blocks_write.argtypes = [_vp, _i, _i, _u64, _i, ctypes.POINTER(_u64)]
lib.moira_blocks_write_gz.argtypes = [_vp, _i, _i, _u64, _i, _i, ctypes.POINTER(_u64)]
lib.moira_gz_scan.argtypes = [_vp, _u64, ctypes.POINTER(_u64), ctypes.POINTER(_u64)]
lib.moira_gz_inflate.argtypes = [_vp, _u64, _i, ctypes.POINTER(_vp), ctypes.POINTER(_u64)]
lib.moira_gz_free.argtypes = [_vp]
lib.moira_gz_deflate.argtypes = [_vp, _u64, _i, _i, _i, _u64, ctypes.POINTER(_u64)]
lib.moira_gz_eof.argtypes = [_i, _u64, ctypes.POINTER(_u64)]
lib.moira_blocks_recycle.argtypes = [_vp]
lib.moira_blocks_free.argtypes = [_vp]
lib.moira_fastq_headers.argtypes = [_vp, _u64, _vp, _u64, _i, _vp, _vp]
lib.moira_fastq_split.argtypes = [_vp, _u64, _i, _vp]
lib.moira_collapse.argtypes = [_vp, _vp, _vp, _vp, _u64, _i, _vp, ctypes.POINTER(_u64), _vp, _vp, _vp, _vp, _vp]
lib.moira_set_host_threads.argtypes = [_i]
lib.moira_fp64_peak.argtypes = [_vp, _i, ctypes.POINTER(_dbl), ctypes.POINTER(_dbl)]
lib.moira_ctx_launch_count.argtypes = [_vp, ctypes.POINTER(_u64)]
lib.moira_comm_unique_id.argtypes = [_vp]
lib.moira_comm_init.argtypes = [_vp, _vp, _i, _i]
lib.moira_comm_init_all.argtypes = [ctypes.POINTER(_vp), _i]
lib.moira_comm_info.argtypes = [_vp, ctypes.POINTER(_i), ctypes.POINTER(_i)]
lib.moira_reduce_counters_device.argtypes = [_vp, _vp, _vp]
lib.moira_reduce_counters.argtypes = [_vp, _vp]
lib.moira_reduce_counters_all.argtypes = [ctypes.POINTER(_vp), _i, ctypes.POINTER(_vp)]
lib.moira_link_probe.argtypes = [_vp, _u64, _i, ctypes.POINTER(_dbl), ctypes.POINTER(_dbl)]
lib.moira_ctx_set_timing.argtypes = [_vp, _i]
lib.moira_ctx_last_contig_ms.argtypes = [_vp, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(_i)]
lib.moira_ctx_last_kernel_ms.argtypes = [_vp, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_char_p)]
_cpp = ctypes.POINTER(ContigParams)
lib.moira_contig_params_default.restype = None
lib.moira_contig_params_default.argtypes = [_cpp]
lib.moira_filter_pairs.argtypes = [_vp, _vp, _u64, _vp, _u64, _vp, _vp, _vp, _vp, _u64, _vp, _u64, _vp, _vp, _vp, _i, _u64, _cpp, _i, _pp, _u64,
                                   _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]
lib.moira_nw_align.argtypes = [_vp, ctypes.c_char_p, ctypes.c_char_p, _i, _i, _i, ctypes.c_char_p, ctypes.c_char_p,
                               ctypes.POINTER(_u64), ctypes.POINTER(ctypes.c_int64)]
lib.moira_make_contig.argtypes = [_vp, ctypes.c_char_p, _vp, _u64, ctypes.c_char_p, _vp, _u64, _cpp, ctypes.c_char_p, _vp,
                                  ctypes.POINTER(_u64), ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32),
                                  ctypes.POINTER(ctypes.c_int32)]
for _name in EXPORTS:
    if _name not in ("moira_last_error", "moira_params_default", "moira_contig_params_default"):
        getattr(lib, _name).restype = _i

if lib.moira_abi_version() != ABI_VERSION:
    raise ImportError("libmoira_b200.so ABI %d != binding ABI %d" % (lib.moira_abi_version(), ABI_VERSION))


def last_error() -> str:
    return (lib.moira_last_error() or b"").decode("utf-8", "replace")


def check(rc: int) -> None:
    if rc != OK:
        raise MoiraError(rc, last_error())
