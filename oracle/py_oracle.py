"""oracle/py_oracle.py -- TEST INFRASTRUCTURE ONLY (never imported by moira_b200/).

Python-3 CPU restatement of the parts of moira's filtering path that exist only as Python in
the reference (all citations relative to /root/reference/moira/moira.py, which is Python 2
and cannot be imported here):

  * interpolate                     moira.py:1723-1733
  * calculate_errors_poisson        moira.py:1637-1679
  * expected_error (Lambda only)    moira.py:1654-1663 (north_star's added "expected_error" mode)
  * filter half of process_data     moira.py:806-833
  * accept/reject of write_results  moira.py:872-883, 911-922, 925-946, 949-970
  * collapse + abundance sort       moira.py:459-475, 491-504
  * fastq / fasta+qual parsing      moira.py:1093-1204 (record semantics only)

plus thin ctypes loaders for the C restatement (oracle/liboracle.so, pb_oracle.c) and for the
unmodified reference binary (oracle/_ref/bernoulli.so, ref_shim.c).

Parity status: PINNED -- tests/test_oracle.py checks these functions against the reference's
known-answer tests (test/test_moira.py:39-45, 63-70), the golden partitions under
test/test_results/ and outputs of the compiled reference committed in tests/golden/.
Python floats are IEEE doubles and `**`, math.exp, math.factorial and int/float division are
the very operations the reference executes, so the arithmetic is the reference's own.
"""
from __future__ import annotations

import ctypes
import importlib.util
import math
import os
from collections import OrderedDict

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(_HERE, "liboracle.so")
REF_SO = os.path.join(_HERE, "_ref", "bernoulli.so")

N_MARK = 0xFF   # 'N'  (in-band slab encoding, include/moira_b200.h)
n_MARK = 0xFE   # 'n'
PAD = 0xFD


# --------------------------------------------------------------------------------------------
# loaders
# --------------------------------------------------------------------------------------------
_oracle_lib = None
_ref_mod = None
_ref_lib = None


def oracle_lib():
    """ctypes handle on oracle/liboracle.so (C restatement)."""
    global _oracle_lib
    if _oracle_lib is None:
        lib = ctypes.CDLL(ORACLE_SO)
        lib.oracle_error_prob.restype = ctypes.c_double
        lib.oracle_error_prob.argtypes = [ctypes.c_int]
        lib.oracle_binomial_pmf.restype = ctypes.c_double
        lib.oracle_binomial_pmf.argtypes = [ctypes.c_double, ctypes.c_int, ctypes.c_int]
        lib.oracle_interpolate.restype = ctypes.c_double
        lib.oracle_interpolate.argtypes = [ctypes.c_int, ctypes.c_double, ctypes.c_int,
                                           ctypes.c_double, ctypes.c_double]
        for name in ("oracle_pb", "oracle_pb_faithful"):
            f = getattr(lib, name)
            f.restype = ctypes.c_int
            f.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_int), ctypes.c_int,
                          ctypes.c_double, ctypes.POINTER(ctypes.c_double),
                          ctypes.POINTER(ctypes.c_int)]
        lib.oracle_pb_batch.restype = ctypes.c_int
        lib.oracle_pb_batch.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                        ctypes.c_uint64, ctypes.c_double, ctypes.c_int,
                                        ctypes.c_void_p, ctypes.c_void_p]
        lib.oracle_tables.restype = None
        lib.oracle_tables.argtypes = [ctypes.c_void_p] * 3
        _oracle_lib = lib
    return _oracle_lib


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def ref_module():
    """The unmodified reference extension module (bernoullimodule.c:66-125) under python 3."""
    global _ref_mod
    if _ref_mod is None:
        spec = importlib.util.spec_from_file_location("bernoulli", REF_SO)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _ref_mod = mod
    return _ref_mod


def ref_lib():
    global _ref_lib
    if _ref_lib is None:
        ref_module()  # makes sure libpython symbols are resolved in-process first
        lib = ctypes.PyDLL(REF_SO)
        lib.ref_pb_batch.restype = ctypes.c_int
        lib.ref_pb_batch.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                     ctypes.c_uint64, ctypes.c_double, ctypes.c_void_p,
                                     ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64]
        _ref_lib = lib
    return _ref_lib


# --------------------------------------------------------------------------------------------
# single-read calculators
# --------------------------------------------------------------------------------------------
def pb_c(contig: str, quals, alpha: float, faithful: bool = False):
    """C restatement of bernoulli.calculate_errors_PB incl. the binding's Q==0 -> 1 remap
    (bernoullimodule.c:104-107).  Returns (ee, Ns)."""
    lib = oracle_lib()
    q = [1 if int(v) == 0 else int(v) for v in quals]
    arr = (ctypes.c_int * max(1, len(q)))(*q)
    ee = ctypes.c_double()
    ns = ctypes.c_int()
    fn = lib.oracle_pb_faithful if faithful else lib.oracle_pb
    rc = fn(contig.encode("latin-1"), arr, len(q), float(alpha), ctypes.byref(ee), ctypes.byref(ns))
    if rc != 0:
        raise RuntimeError("oracle_pb failed rc=%d" % rc)
    return ee.value, ns.value


def interpolate(errors1, prob1, errors2, prob2, alpha):
    """moira.py:1723-1733."""
    result = errors1 + ((errors2 - errors1) * ((1 - alpha) - prob1) / (prob2 - prob1))
    if result < 0:
        result = 0
    return result


def calculate_errors_poisson(sequence, quals, alpha):
    """moira.py:1637-1679.  Only uppercase 'N' is skipped (moira.py:1660)."""
    sequence = str(sequence)
    quals = [int(q) for q in quals]
    alpha = float(alpha)
    if len(sequence) != len(quals):
        raise ValueError("length mismatch")
    if alpha <= 0 or alpha > 1:
        raise ValueError("Alpha must be between 0 (not included) and 1.")
    Lambda = 0
    Ns = 0
    for base, qscore in zip(sequence, quals):
        if qscore < 0:
            raise ValueError("Qualities must have positive values.")
        if base == "N":
            Ns += 1
        else:
            Lambda += 10 ** (qscore / -10.0)
    accumulated = [0]
    expected_errors = 0
    while 1:
        probability = (math.exp(-Lambda) * (Lambda ** expected_errors)) / (math.factorial(expected_errors))
        accumulated.append(accumulated[-1] + probability)
        if accumulated[-1] > (1 - alpha):
            break
        expected_errors += 1
    expected_errors = interpolate(expected_errors - 1, accumulated[-2], expected_errors,
                                  accumulated[-1], alpha)
    return expected_errors, Ns


def expected_error(sequence, quals):
    """Lambda of moira.py:1654-1663: sequential sum of p_i over non-'N' bases."""
    Lambda = 0.0
    Ns = 0
    for base, qscore in zip(sequence, quals):
        if base == "N":
            Ns += 1
        else:
            Lambda += 10 ** (int(qscore) / -10.0)
    return Lambda, Ns


# --------------------------------------------------------------------------------------------
# process_data (filter half) and the decision of write_results
# --------------------------------------------------------------------------------------------
class Args:
    """Defaults of the reference CLI (moira.py:649-668) / its test-suite (test_moira.py:130-135)."""

    def __init__(self, **kw):
        self.alpha = 0.005
        self.uncert = 0.01
        self.maxerrors = None
        self.error_calc = "poisson_binomial"
        self.ambigs = "treat_as_errors"
        self.round = False
        self.truncate = None
        self.collapse = True
        self.__dict__.update(kw)


def process_filter(contig, quals, args, pb=pb_c):
    """moira.py:806-833 for an already-assembled contig.  Returns (contig, quals, ee)."""
    if args.truncate:
        contig, quals = contig[:args.truncate], quals[:args.truncate]          # :806-807
    quals = [q if q > 0 else 1 for q in quals]                                 # :814
    if args.error_calc in ("poisson_binomial", "poisson_binomial_py"):
        ee, ns = pb(contig, quals, args.alpha)                                 # :817
    elif args.error_calc == "poisson":
        ee, ns = calculate_errors_poisson(contig, quals, args.alpha)           # :823
    elif args.error_calc == "expected_error":
        ee, ns = expected_error(contig, quals)
    else:
        raise ValueError(args.error_calc)
    if args.ambigs == "treat_as_errors":
        ee = ee + ns                                                           # :827-828
    if args.round:
        ee = math.floor(ee)                                                    # :830-831
    return contig, quals, ee


REASON_NONE, REASON_ERRORS, REASON_LENGTH, REASON_AMBIGS = 0, 1, 2, 3


def decide(sequence, ee, args):
    """Accept/reject + reason, in write_results' precedence (moira.py:872-970), single-end."""
    if args.truncate and len(sequence) < args.truncate:                        # :872
        return False, REASON_LENGTH
    if "N" in sequence and args.ambigs == "disallow":                          # :911
        return False, REASON_AMBIGS
    if args.maxerrors:                                                         # :925
        ok = ee <= args.maxerrors                                              # :926
    else:
        ok = ee <= len(sequence) * args.uncert                                 # :950
    return (True, REASON_NONE) if ok else (False, REASON_ERRORS)


def collapse_and_decide(records, args, pb=pb_c):
    """records: iterable of (header, seq, quals).  Mirrors the main loop (moira.py:455-475),
    the epilogue (moira.py:491-504) and the decision.  Returns (good, bad) where each maps the
    representative header -> list of member headers (the .names file content, :938/:955)."""
    uniques = OrderedDict()
    for header, seq, quals in records:
        contig, cq, ee = process_filter(seq, quals, args, pb=pb)
        u = uniques.get(contig)
        if u is None:
            uniques[contig] = {"rep_header": header, "rep_errors": ee, "names": [header]}
        elif ee < u["rep_errors"]:                                             # :466
            u["rep_header"] = header
            u["rep_errors"] = ee
            u["names"].insert(0, header)                                       # :470
        else:
            u["names"].append(header)
    good, bad = OrderedDict(), OrderedDict()
    order = sorted(uniques, key=lambda s: len(uniques[s]["names"]), reverse=True)   # :492
    for seq in order:
        u = uniques[seq]
        ok, reason = decide(seq, u["rep_errors"], args)
        (good if ok else bad)[u["rep_header"]] = (list(u["names"]), u["rep_errors"], reason)
    return good, bad


# --------------------------------------------------------------------------------------------
# record parsing (semantics of moira.py:1093-1204, single-end)
# --------------------------------------------------------------------------------------------
def _norm_header(line, lead):
    return line.strip().replace("\t", " ").split(" ")[0].lstrip(lead).replace(":", "_")


def parse_fastq_text(text, offset=33):
    lines = text.splitlines()
    out = []
    for i in range(0, len(lines) - 3, 4):
        header = _norm_header(lines[i], "@")                                   # :1175
        seq = lines[i + 1].strip()
        quals = [ord(c) - offset for c in lines[i + 3].strip()]                # :1177
        out.append((header, seq, quals))
    return out


def parse_fasta_qual_text(fasta_text, qual_text):
    f = fasta_text.splitlines()
    q = qual_text.splitlines()
    out = []
    for i in range(0, len(f) - 1, 2):
        header = _norm_header(f[i], ">")                                       # :1121
        seq = f[i + 1].strip()
        quals = [int(x) for x in q[i + 1].strip().replace("\t", " ").split(" ")]   # :1124
        out.append((header, seq, quals))
    return out


# --------------------------------------------------------------------------------------------
# packed-slab helpers (in-band encoding of include/moira_b200.h) and batch drivers
# --------------------------------------------------------------------------------------------
def pack_records(seqs, quals_list, lower_n_ambiguous=True, align=16):
    """Pack reads into one uint8 slab: byte = Phred score, 0xFF for 'N', 0xFE for 'n'
    (only when lower_n_ambiguous -- the C core skips both, bernoullimodule.c:196; the Python
    calculators only 'N', moira.py:1605/1660).  Rows start 16-byte aligned, padding = 0xFD."""
    n = len(seqs)
    lengths = np.array([len(s) for s in seqs], dtype=np.uint32)
    strides = (lengths.astype(np.uint64) + (align - 1)) // align * align
    offsets = np.zeros(n, dtype=np.uint64)
    if n:
        offsets[1:] = np.cumsum(strides)[:-1]
    total = int(strides.sum()) if n else 0
    slab = np.full(max(total, align), PAD, dtype=np.uint8)
    for i, (s, q) in enumerate(zip(seqs, quals_list)):
        row = np.asarray(q, dtype=np.int64)
        if row.size and (row.min() < 0 or row.max() > 0xFC):
            raise ValueError("quality outside 0..252")
        row = row.astype(np.uint8)
        sb = np.frombuffer(s.encode("latin-1"), dtype=np.uint8)
        row = np.where(sb == ord("N"), np.uint8(N_MARK), row)
        if lower_n_ambiguous:
            row = np.where(sb == ord("n"), np.uint8(n_MARK), row)
        o = int(offsets[i])
        slab[o:o + len(s)] = row
    return slab, offsets, lengths


def pb_batch(slab, offsets, lengths, alpha, faithful=False):
    """C restatement over a packed slab -> (ee float64[n], ns int32[n])."""
    lib = oracle_lib()
    n = len(lengths)
    ee = np.zeros(n, dtype=np.float64)
    ns = np.zeros(n, dtype=np.int32)
    slab = np.ascontiguousarray(slab, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    lengths = np.ascontiguousarray(lengths, dtype=np.uint32)
    rc = lib.oracle_pb_batch(slab.ctypes.data, offsets.ctypes.data, lengths.ctypes.data, n,
                             float(alpha), int(faithful), ee.ctypes.data, ns.ctypes.data)
    if rc != 0:
        raise RuntimeError("oracle_pb_batch rc=%d" % rc)
    return ee, ns


def ref_batch(slab, offsets, lengths, alpha, n_threads=1, stack_bytes=256 << 20):
    """The unmodified reference's test() over a packed slab, on big-stack worker threads."""
    lib = ref_lib()
    n = len(lengths)
    ee = np.zeros(n, dtype=np.float64)
    ns = np.zeros(n, dtype=np.int32)
    slab = np.ascontiguousarray(slab, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    lengths = np.ascontiguousarray(lengths, dtype=np.uint32)
    rc = lib.ref_pb_batch(slab.ctypes.data, offsets.ctypes.data, lengths.ctypes.data, n,
                          float(alpha), ee.ctypes.data, ns.ctypes.data, int(n_threads),
                          int(stack_bytes))
    if rc != 0:
        raise RuntimeError("ref_pb_batch rc=%d" % rc)
    return ee, ns


def poisson_batch(slab, offsets, lengths, alpha):
    """Reference Poisson arithmetic over a packed slab (pure Python; small cases only)."""
    n = len(lengths)
    ee = np.zeros(n, dtype=np.float64)
    ns = np.zeros(n, dtype=np.int32)
    lam = np.zeros(n, dtype=np.float64)
    for r in range(n):
        row = slab[int(offsets[r]):int(offsets[r]) + int(lengths[r])]
        seq = "".join("N" if b >= n_MARK else "A" for b in row.tolist())
        q = [2 if b >= n_MARK else (1 if b == 0 else b) for b in row.tolist()]
        ee[r], ns[r] = calculate_errors_poisson(seq, q, alpha)
        lam[r], _ = expected_error(seq, q)
    return ee, ns, lam


def decide_batch(ee_raw, ns, lengths, has_N, *, thr_kind, thr, ambigs, round_flag, truncate):
    """Vectorised A10/A11 (moira.py:827-831, 872-970) on raw (ee, Ns).  lengths are the
    ORIGINAL read lengths; returns (accept bool[n], reason uint8[n], ee_final float64[n])."""
    ee = np.asarray(ee_raw, dtype=np.float64).copy()
    if ambigs == "treat_as_errors":
        ee = ee + np.asarray(ns, dtype=np.float64)
    if round_flag:
        ee = np.floor(ee)
    lengths = np.asarray(lengths, dtype=np.int64)
    eff = np.minimum(lengths, truncate) if truncate else lengths
    if thr_kind == "maxerrors":
        ok = ee <= thr
    else:
        ok = ee <= eff.astype(np.float64) * thr
    reason = np.where(ok, REASON_NONE, REASON_ERRORS).astype(np.uint8)
    if ambigs == "disallow":
        amb = np.asarray(has_N, dtype=bool)
        reason = np.where(amb, REASON_AMBIGS, reason).astype(np.uint8)
        ok = ok & ~amb
    if truncate:
        short = lengths < truncate
        reason = np.where(short, REASON_LENGTH, reason).astype(np.uint8)
        ok = ok & ~short
    return ok, reason, ee


# --------------------------------------------------------------------------------------------
# paired-end contig constructor: wrappers over oracle/contig_oracle.c and the
# unmodified reference aligner oracle/_ref/nw_align.so
# --------------------------------------------------------------------------------------------
REF_NW_SO = os.path.join(_HERE, "_ref", "nw_align.so")
CONSENSUS = {"best": 0, "sum": 1, "posterior": 2}
_ref_nw = None
_contig_ready = False


def have_ref_nw() -> bool:
    return os.path.exists(REF_NW_SO)


def ref_nw_module():
    """The unmodified reference aligner (moira/nw_align.pyx), cythonised by oracle/Makefile."""
    global _ref_nw
    if _ref_nw is None:
        spec = importlib.util.spec_from_file_location("nw_align", REF_NW_SO)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _ref_nw = mod
    return _ref_nw


def _contig_lib():
    global _contig_ready
    lib = oracle_lib()
    if not _contig_ready:
        i32p = ctypes.POINTER(ctypes.c_int32)
        ip = ctypes.POINTER(ctypes.c_int)
        lib.oracle_reverse_complement.restype = ctypes.c_int
        lib.oracle_reverse_complement.argtypes = [ctypes.c_char_p, i32p, ctypes.c_int, ctypes.c_char_p, i32p]
        lib.oracle_nw_align.restype = ctypes.c_int
        lib.oracle_nw_align.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_int, ctypes.c_char_p, ctypes.c_char_p, ip, ctypes.POINTER(ctypes.c_long)]
        lib.oracle_make_contig.restype = ctypes.c_int
        lib.oracle_make_contig.argtypes = [ctypes.c_char_p, i32p, ctypes.c_int, ctypes.c_char_p, i32p, ctypes.c_int, ctypes.c_int,
                                           ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                           ctypes.c_char_p, i32p, ip, ip, ip, ip]
        lib.oracle_pair_to_contig.restype = ctypes.c_int
        lib.oracle_pair_to_contig.argtypes = [ctypes.c_char_p, i32p, ctypes.c_int, ctypes.c_char_p, i32p, ctypes.c_int] + \
            [ctypes.c_int] * 8 + [ctypes.c_char_p, i32p, ip, ip, ip, ip]
        _contig_ready = True
    return lib


def _i32(values):
    arr = (ctypes.c_int32 * max(1, len(values)))(*values)
    return arr


def reverse_complement(sequence, quals=None):
    """moira.py:1207-1236: reverse complement of an IUPAC string, qualities reversed along."""
    lib = _contig_lib()
    n = len(sequence)
    out = ctypes.create_string_buffer(n + 1)
    qi = _i32(list(quals)) if quals else None
    qo = (ctypes.c_int32 * max(1, n))()
    if quals and len(quals) != n:
        raise ValueError("LengthMismatchError")
    rc = lib.oracle_reverse_complement(sequence.encode("latin-1"), qi, n, out, qo)
    if rc:
        raise ValueError('"%s" is not a recognizable IUPAC-coded base.' % sequence[rc - 1])
    s = out.raw[:n].decode("latin-1")
    return (s, list(qo[:n])) if quals else s


def nw_align(seq_1, seq_2, match, mismatch, gap):
    """nw_align.pyx:49-145 with refine_overlap=True -> (seq_1_aligned, seq_2_aligned, score)."""
    lib = _contig_lib()
    l1, l2 = len(seq_1), len(seq_2)
    a1 = ctypes.create_string_buffer(l1 + l2 + 1)
    a2 = ctypes.create_string_buffer(l1 + l2 + 1)
    n = ctypes.c_int()
    score = ctypes.c_long()
    if lib.oracle_nw_align(seq_1.encode("latin-1"), l1, seq_2.encode("latin-1"), l2, match, mismatch, gap, a1, a2,
                           ctypes.byref(n), ctypes.byref(score)):
        raise MemoryError
    return a1.raw[:n.value].decode("latin-1"), a2.raw[:n.value].decode("latin-1"), score.value


def make_contig(forward_aligned, forward_quals, reverse_aligned, reverse_quals, insert, deltaq, consensus_qscore, qscore_cap,
                trim_overlap):
    """moira.py:1375-1558 -> (contig, contig_quals, overlap_length, gaps, mismatches)."""
    lib = _contig_lib()
    if consensus_qscore not in CONSENSUS:
        raise ValueError('consensus_qscore must be "best", "sum" or "posterior".')
    n = len(forward_aligned)
    assert len(reverse_aligned) == n
    contig = ctypes.create_string_buffer(n + 1)
    cq = (ctypes.c_int32 * max(1, n))()
    clen, ov, gaps, mism = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    rc = lib.oracle_make_contig(forward_aligned.encode("latin-1"), _i32(list(forward_quals)), len(forward_quals),
                                reverse_aligned.encode("latin-1"), _i32(list(reverse_quals)), len(reverse_quals), n, int(insert),
                                int(deltaq), CONSENSUS[consensus_qscore], int(qscore_cap), int(bool(trim_overlap)), contig, cq,
                                ctypes.byref(clen), ctypes.byref(ov), ctypes.byref(gaps), ctypes.byref(mism))
    if rc == -2:
        raise ValueError("LengthMismatchError")
    if rc:
        raise ValueError("make_contig failed with code %d" % rc)
    return contig.raw[:clen.value].decode("latin-1"), list(cq[:clen.value]), ov.value, gaps.value, mism.value


def pair_to_contig(forward_sequence, forward_quals, reverse_sequence, reverse_quals, match=1, mismatch=-1, gap=-2, insert=20,
                   deltaq=6, consensus_qscore="best", qscore_cap=40, trim_overlap=False):
    """The paired branch of process_data (moira.py:791-803) for one pair of raw reads."""
    lib = _contig_lib()
    l1, l2 = len(forward_sequence), len(reverse_sequence)
    contig = ctypes.create_string_buffer(l1 + l2 + 1)
    cq = (ctypes.c_int32 * max(1, l1 + l2))()
    clen, ov, gaps, mism = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    rc = lib.oracle_pair_to_contig(forward_sequence.encode("latin-1"), _i32(list(forward_quals)), l1,
                                   reverse_sequence.encode("latin-1"), _i32(list(reverse_quals)), l2, match, mismatch, gap, insert,
                                   deltaq, CONSENSUS[consensus_qscore], qscore_cap, int(bool(trim_overlap)), contig, cq,
                                   ctypes.byref(clen), ctypes.byref(ov), ctypes.byref(gaps), ctypes.byref(mism))
    if rc:
        raise ValueError("pair_to_contig failed with code %d" % rc)
    return contig.raw[:clen.value].decode("latin-1"), list(cq[:clen.value]), ov.value, gaps.value, mism.value
