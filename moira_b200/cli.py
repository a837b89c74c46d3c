"""moira_b200 command line: the Python-3 host of the B200 path.

Keeps the reference CLI's flags, defaults, input sniffing, collapse logic, output files and
formatting (fpusan/moira v1.3.2, moira/moira.py: parse_arguments :581-675, check_arguments :678-781,
main :264-578, write_results :842-970, parsers :1058-1204), but instead of calling
`bernoulli.calculate_errors_PB` once per read (moira.py:817) it parses reads in batches straight into
quality slabs and runs them through libmoira_b200.so (one moira_filter_batch per batch).

Flows: single-end FASTQ (text -> device parser -> filter -> device dereplication, sharded over --devices), single-end
FASTA + QUAL (native host parser), and --paired / --only_contig (contigs built on the device).  Records
are never turned into Python objects: the parsers leave byte ranges, the device leaves decisions, the native writer
(moira_format_records) turns both into the output files.
`--error_calc` adds `expected_error` (north_star) and drops `bootstrap` (deprecated, moira.py:772-776);
`poisson_binomial_py` is an alias of `poisson_binomial` with the Python twin's 'N'-only rule.
"""
from __future__ import annotations

import argparse
import bz2
import gzip
import io
import mmap
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import _lib as L
from .api import ContigParams, Context, FilterParams, MoiraError, collapse, parse_fasta_qual, parse_fastq
from .contig import LengthMismatchError  # noqa: F401  (moira.py:1024-1038; one class for the parsers, the calculators and make_contig)

__version__ = "0.1.0 (moira 1.3.2 compatible)"

BATCH_BYTES = 64 << 20   # FASTQ text per batch


# ---- exceptions of the reference (moira.py:973-1055) ---------------------------------------------
class ReturnedNaNError(Exception):
    pass


class UnpairedFilesError(Exception):
    pass


class NameMismatchError(Exception):
    def __init__(self, *headers):
        super().__init__("Sequence headers do not match: %s" % ", ".join(str(h) for h in headers))


class EmptySeqError(Exception):
    def __init__(self, header, fname):
        super().__init__("Empty sequence for %s in %s" % (header, fname))


class EmptyQualError(Exception):
    def __init__(self, header, fname):
        super().__init__("Empty qualities for %s in %s" % (header, fname))


# ---- arguments (moira.py:581-781) ------------------------------------------------------------------
def parse_arguments(argv=None):
    def str2bool(value):
        return value.lower() in ("yes", "true", "t", "1")

    parser = argparse.ArgumentParser(description="Perform quality filtering on a set of sequences (B200 path).")
    general = parser.add_argument_group("General options")
    general.add_argument("-ff", "--forward_fasta", type=str)
    general.add_argument("-fq", "--forward_qual", type=str)
    general.add_argument("-rf", "--reverse_fasta", type=str)
    general.add_argument("-rq", "--reverse_qual", type=str)
    general.add_argument("-ffq", "--forward_fastq", type=str)
    general.add_argument("-rfq", "--reverse_fastq", type=str)
    general.add_argument("-l", "--relabel", type=str)
    general.add_argument("-o", "--output_format", type=str, default="fasta", choices=("fasta", "fastq"))
    general.add_argument("-pi", "--pipeline", type=str, default="mothur", choices=("mothur", "USEARCH"))
    general.add_argument("-op", "--output_prefix", type=str)
    general.add_argument("-oc", "--output_compression", type=str, default="none", choices=("none", "gz", "bz2"))
    general.add_argument("-p", "--processors", type=int, default=1,
                         help="Accepted for compatibility; the GPU path does not use worker processes (see --devices).")
    general.add_argument("--paired", action="store_true")
    general.add_argument("-fo", "--fastq_offset", type=int, default=33)
    general.add_argument("--only_contig", action="store_true")
    general.add_argument("--silent", action="store_true")
    general.add_argument("--nowarnings", action="store_true")
    general.add_argument("--doc", action="store_true")
    general.add_argument("--device", type=int, default=0, help="CUDA device index (B200 path only).")
    general.add_argument("--devices", type=str, default=None,
                         help="'all' or a comma-separated list of CUDA devices: reads are sharded by contiguous chunk, one context "
                              "and one host thread per GPU (the B200 counterpart of --processors); overrides --device.  'all' takes "
                              "one GPU per 2 GiB of input text, at most every GPU of the box; a list is taken as given.")
    constructor = parser.add_argument_group("Contig construction options")     # moira.py:629-646
    constructor.add_argument("-m", "--match", type=int, default=1)
    constructor.add_argument("-x", "--mismatch", type=int, default=-1)
    constructor.add_argument("-g", "--gap", type=int, default=-2)
    constructor.add_argument("--trim_overlap", action="store_true")
    constructor.add_argument("-i", "--insert", type=int, default=20)
    constructor.add_argument("-d", "--deltaq", type=int, default=6)
    constructor.add_argument("-q", "--consensus_qscore", type=str, default="best", choices=("best", "sum", "posterior"))
    constructor.add_argument("-z", "--qscore_cap", type=int, default=40)
    filtering = parser.add_argument_group("Sequence filtering options")
    filtering.add_argument("-c", "--collapse", type=str2bool, default="True")
    filtering.add_argument("-t", "--truncate", type=int)
    filtering.add_argument("-mo", "--min_overlap", type=int)
    filtering.add_argument("-e", "--error_calc", type=str, default="poisson_binomial",
                           choices=("poisson_binomial", "poisson_binomial_py", "poisson", "expected_error"))
    filtering.add_argument("-n", "--ambigs", type=str, default="treat_as_errors",
                           choices=("disallow", "ignore", "treat_as_errors"))
    filtering.add_argument("-r", "--round", action="store_true")
    err_uncert = filtering.add_mutually_exclusive_group()
    err_uncert.add_argument("-u", "--uncert", type=float, default=0.01)
    err_uncert.add_argument("-me", "--maxerrors", type=float)
    filtering.add_argument("-a", "--alpha", type=float, default=0.005)
    args = parser.parse_args(argv)
    if isinstance(args.collapse, str):
        args.collapse = str2bool(args.collapse)
    return args


def check_arguments(args, out=sys.stdout):
    """moira.py:678-781 for the flags this build implements."""
    ok = True

    def warn(msg):
        if not args.nowarnings:
            print(msg, file=out)

    if args.doc:
        print(__doc__, file=out)
        return False
    if args.only_contig:                                                       # moira.py:695-696
        args.paired = True
    if not args.forward_fastq and (not args.forward_fasta or not args.forward_qual):
        warn("- You must at least provide one fastq file, or a fasta and quality files.")
        ok = False
    if args.paired and not args.reverse_fastq and (not args.reverse_fasta or not args.reverse_qual):
        warn("- You must provide one reverse fastq file, or reverse fasta and quality files.")
        ok = False
    if args.match < 0:
        warn("- Needleman-Wunsch match score must be a non-negative integer.")
        ok = False
    if args.mismatch > 0:
        warn("- Needleman-Wunsch mismatch penalty must be a non-positive integer.")
        ok = False
    if args.gap > 0:
        warn("- Needleman-Wunsch gap penalty must be a non-positive integer.")
        ok = False
    if args.insert < 1:
        warn("- The contig constructor insert parameter must be a positive integer.")
        ok = False
    if args.deltaq < 1:
        warn("- The contig constructor deltaq parameter must be a positive integer.")
        ok = False
    if args.qscore_cap < 0:
        warn("- The contig constructor qscore_cap parameter must be a non-negative integer.")
        ok = False
    if args.min_overlap is not None and args.min_overlap <= 0:
        warn("- The min_overlap parameter must be greater than 0.")
        ok = False
    if not 0 < args.uncert <= 1:
        warn("- The uncert parameter must be between 0 (not included) and 1.")
        ok = False
    if args.maxerrors is not None and args.maxerrors <= 0:
        warn("- The maxerrors parameter must be greater than 0.")
        ok = False
    if not 0 < args.alpha < 1:
        warn("- The alpha parameter must be between 0 (not included) and 1.")
        ok = False
    if args.truncate and args.truncate <= 0:
        warn("- The truncate parameter must be greater than 0.")
        ok = False
    if not ok:
        warn("\nFor more info type moira.py -h or moira.py --doc.\n")
        return False
    if (args.reverse_fasta or args.reverse_fastq) and not args.paired:
        warn("You provided a reverse sequence file, but not the --paired flag. Note that only the forward file will be processed.\n")
    if args.min_overlap and not args.paired:
        warn("You specified a value for --min_overlap, but not the --paired flag. Note that contigs will not be assembled.")
    return True


# ---- input (moira.py:1058-1204) -------------------------------------------------------------------
def open_input(filename):
    """Sniff gzip / bzip2 by magic bytes (moira.py:1065-1068) and return a binary file object."""
    fh = io.open(filename, mode="rb", buffering=1 << 20)
    start = fh.peek(3)[:3]
    if start.startswith(b"\x1f\x8b\x08"):
        fh.close()
        return gzip.GzipFile(filename=filename)
    if start.startswith(b"\x42\x5a\x68"):
        fh.close()
        return bz2.BZ2File(filename)
    return fh


def _norm_header(raw: str, lead: str) -> str:
    return raw.strip().replace("\t", " ").split(" ")[0].lstrip(lead).replace(":", "_")


def load_text(filename):
    """Whole input file as a uint8 array: plain files are memory-mapped (no copy; the library stages pageable text through
    its own pinned ring), gzip / bzip2 files (sniffed by magic bytes, moira.py:1065-1068) are inflated into memory."""
    with io.open(filename, mode="rb") as fh:
        start = fh.read(3)
        if start.startswith(b"\x1f\x8b\x08"):
            # moira_gz_inflate: blocked gzip (BGZF: bgzip, Illumina's FASTQ writers) on all host threads, any other gzip
            # file on one -- straight from the mapped file into memory the library owns
            from .api import gz_inflate
            mm = mmap.mmap(fh.fileno(), 0, access=mmap.ACCESS_READ)
            try:
                text = gz_inflate(np.frombuffer(mm, dtype=np.uint8))
            except MoiraError as exc:
                if exc.code != L.ERR_PARSE:
                    raise
                raise IOError("%s: %s" % (filename, exc.message)) from None
            return text, text
        elif start.startswith(b"\x42\x5a\x68"):
            data = bz2.BZ2File(filename).read()
        else:
            size = os.fstat(fh.fileno()).st_size
            if size == 0:
                return np.zeros(0, np.uint8), None
            mm = mmap.mmap(fh.fileno(), 0, access=mmap.ACCESS_READ)
            return np.frombuffer(mm, dtype=np.uint8), mm
    return np.frombuffer(data, dtype=np.uint8), data


def _raise_parse_error(exc, *fnames):
    """MOIRA_ERR_PARSE -> the reference's exception classes (moira.py:994-1055)."""
    if exc.code != L.ERR_PARSE:
        raise exc
    name = exc.message.split(":")[0]
    if name == "NameMismatchError":
        raise NameMismatchError(exc.message, "") from None
    if name == "EmptySeqError":
        raise EmptySeqError(exc.message, fnames[0]) from None
    if name == "EmptyQualError":
        raise EmptyQualError(exc.message, fnames[-1]) from None
    raise LengthMismatchError(exc.message, *fnames) from None


def _gather(buf, off, ln):
    """Concatenation of buf[off[i] : off[i] + ln[i]] over i (vectorised)."""
    ln = ln.astype(np.int64)
    total = int(ln.sum())
    if total == 0:
        return np.zeros(0, np.uint8)
    starts = np.cumsum(ln) - ln
    idx = np.repeat(off.astype(np.int64) - starts, ln) + np.arange(total, dtype=np.int64)
    return buf[idx]


def _cut_lines(data: bytes, n_lines: int) -> int:
    """Byte position just after the n_lines-th newline of data (len(data) if it has fewer)."""
    idx = np.flatnonzero(np.frombuffer(data, dtype=np.uint8) == 10)
    return int(idx[n_lines - 1]) + 1 if 0 < n_lines <= len(idx) else len(data)


def read_fasta_qual_batches(fasta_name, qual_name, lower_n_ambiguous):
    """moira.py:1093-1149, single-end: sequences and qualities on one line each.  Both files are memory-mapped (gz / bz2:
    inflated once), cut at the same record numbers by moira_line_offsets, and every pair of blocks goes through the native
    parser (moira_parse_fasta_qual) as a view -- no text passes through Python."""
    from .api import line_offsets
    ftext, fkeep = load_text(fasta_name)
    qtext, qkeep = load_text(qual_name)
    f_empty = ftext.size == 0 or (ftext.size < 4096 and not ftext.tobytes().strip())
    if f_empty:
        if qtext.size and (qtext.size >= 4096 or qtext.tobytes().strip()):
            raise NameMismatchError("", _norm_header(qtext[:4096].tobytes().decode("latin-1").splitlines()[0], ">"))
        return
    head, qhead = ftext[:1 << 20], qtext[:1 << 20]
    rec_bytes = max(8.0, head.size / max(1.0, np.count_nonzero(head == 10) / 2.0)) + \
        max(8.0, qhead.size / max(1.0, np.count_nonzero(qhead == 10) / 2.0))
    per_block = max(1, int(2 * BATCH_BYTES / rec_bytes))
    _, nl_f = line_offsets(ftext, np.zeros(0, np.uint64))
    n_blocks = max(1, -(-((nl_f + 1) // 2) // per_block))
    marks = np.arange(1, n_blocks, dtype=np.uint64) * np.uint64(2 * per_block)
    fcut = np.concatenate([[0], line_offsets(ftext, marks)[0], [ftext.size]]).astype(np.int64)
    qcut = np.concatenate([[0], line_offsets(qtext, marks)[0], [qtext.size]]).astype(np.int64)
    for k in range(n_blocks):
        fb, qb = ftext[fcut[k]:fcut[k + 1]], qtext[qcut[k]:qcut[k + 1]]
        if fb.size == 0 and qb.size == 0:
            continue
        try:
            slab, qslab, offsets, lengths, hoff, hlen, soff = parse_fasta_qual(fb, qb, lower_n_ambiguous)
        except MoiraError as exc:
            if exc.code == L.ERR_PARSE:
                name = exc.message.split(":")[0]
                if name == "NameMismatchError":
                    raise NameMismatchError(exc.message, "") from None
                if name == "EmptySeqError":
                    raise EmptySeqError(exc.message, fasta_name) from None
                if name == "EmptyQualError":
                    raise EmptyQualError(exc.message, qual_name) from None
                if name == "LengthMismatchError":
                    raise LengthMismatchError(exc.message, fasta_name, qual_name) from None
                raise ValueError(exc.message) from None
            raise
        if len(lengths) == 0:
            continue
        # arrays only: (fasta text, header offsets / lengths, sequence offsets, plain qualities at the slab's offsets, slab, ...)
        yield fb, hoff, hlen, soff, qslab, slab, offsets, lengths
    del fkeep, qkeep


def _whole_records(fh, carry, lines_per_record, want_lines=None):
    """Next block of whole records of a text file: (text, carry, n_lines, at_eof).  want_lines: exactly that many
    lines (the partner file's block decides); None: whatever a BATCH_BYTES read holds."""
    data = carry
    eof = False
    while True:
        n_lines = data.count(b"\n")
        if want_lines is not None and n_lines >= want_lines:
            break
        block = fh.read(BATCH_BYTES)
        if not block:
            eof = True
            break
        data += block
        if want_lines is None:
            break
    n_lines = data.count(b"\n")
    if eof:
        if data and not data.endswith(b"\n"):
            n_lines += 1
        keep = n_lines - n_lines % lines_per_record if want_lines is None else min(want_lines, n_lines)
        if want_lines is None or keep >= n_lines:
            return data, b"", keep, True
    else:
        keep = n_lines - n_lines % lines_per_record if want_lines is None else want_lines
    pos = _cut_lines(data, keep)
    return data[:pos], data[pos:], keep, eof and pos >= len(data)


def _parse_error_for(exc, fname):
    """MoiraError(ERR_PARSE) of a native parser -> the reference's exception class (moira.py:1178-1183, 1141-1142)."""
    if exc.code != L.ERR_PARSE:
        return exc
    name = exc.message.split(":")[0]
    cls = {"EmptySeqError": EmptySeqError, "EmptyQualError": EmptyQualError}.get(name)
    if cls:
        return cls(exc.message, fname)
    if name == "NameMismatchError":
        return NameMismatchError(exc.message, "")
    return LengthMismatchError(exc.message, fname)


def _check_pair_headers(ftext, fh_off, fh_len, rtext, rh_off, rh_len):
    """Both files must name the same record at every position (moira.py:1199-1200)."""
    fb = ftext if isinstance(ftext, np.ndarray) else np.frombuffer(ftext, dtype=np.uint8)
    rb = rtext if isinstance(rtext, np.ndarray) else np.frombuffer(rtext, dtype=np.uint8)
    if np.array_equal(fh_len, rh_len) and np.array_equal(_gather(fb, fh_off, fh_len), _gather(rb, rh_off, rh_len)):
        return
    for i in range(len(fh_len)):
        fh_ = fb[int(fh_off[i]):int(fh_off[i]) + int(fh_len[i])].tobytes().decode("latin-1").replace(":", "_")
        rh_ = rb[int(rh_off[i]):int(rh_off[i]) + int(rh_len[i])].tobytes().decode("latin-1").replace(":", "_")
        if fh_ != rh_:
            raise NameMismatchError(fh_, None, rh_, None)


def _read_fastq_pair_batches(args, lower_n_ambiguous):
    """Two FASTQ files in lock step without touching the text in Python: both files are memory-mapped (gz / bz2: inflated
    once), moira_line_offsets cuts them at the same record numbers (one parallel newline count per file), and every pair
    of blocks goes through the native parser as a view."""
    from .api import index_fastq, line_offsets
    ftext, fkeep = load_text(args.forward_fastq)
    rtext, rkeep = load_text(args.reverse_fastq)
    if ftext.size == 0 or (ftext.size < 4096 and not ftext.tobytes().strip()):
        if rtext.size and (rtext.size >= 4096 or rtext.tobytes().strip()):
            raise NameMismatchError("", "(the forward file ended first)")
        return
    head = ftext[:1 << 20]
    rec_bytes = max(16.0, head.size / max(1.0, np.count_nonzero(head == 10) / 4.0))
    per_block = max(1, int(BATCH_BYTES / rec_bytes))
    _, nl_f = line_offsets(ftext, np.zeros(0, np.uint64))
    n_blocks = max(1, -(-((nl_f + 3) // 4) // per_block))
    marks = (np.arange(1, n_blocks, dtype=np.uint64) * np.uint64(4 * per_block))
    fcut, _ = line_offsets(ftext, marks)
    rcut, _ = line_offsets(rtext, marks)
    fcut = np.concatenate([[0], fcut, [ftext.size]]).astype(np.int64)
    rcut = np.concatenate([[0], rcut, [rtext.size]]).astype(np.int64)
    for k in range(n_blocks):
        fb, rb = ftext[fcut[k]:fcut[k + 1]], rtext[rcut[k]:rcut[k + 1]]
        if fb.size == 0 and rb.size == 0:
            continue
        parsed = []
        for buf, fname in ((fb, args.forward_fastq), (rb, args.reverse_fastq)):
            try:
                ln, hoff, hlen, soff, qoff = index_fastq(buf)      # the contig kernel reads bases and quality characters from the text itself
            except MoiraError as exc:
                raise _parse_error_for(exc, fname) from None
            parsed.append((buf, hoff, hlen, (buf, buf, soff, qoff, ln, args.fastq_offset)))
        (ft, fh_off, fh_len, fwd), (rt, rh_off, rh_len, rev) = parsed
        n = len(fwd[4])
        if len(rev[4]) != n:
            if n == 0:
                raise NameMismatchError("", "(the forward file ended first)")
            raise NameMismatchError("(%d forward records)" % n, "(%d reverse records)" % len(rev[4]))
        if n == 0:
            continue
        _check_pair_headers(ft, fh_off, fh_len, rt, rh_off, rh_len)
        yield ft, fh_off, fh_len, fwd, rev
    del fkeep, rkeep


def read_pair_batches(args, lower_n_ambiguous):
    """Paired input (moira.py:1093-1204 with both files): blocks of whole records of the forward file and the same
    number of records of the reverse file, parsed natively, headers compared (NameMismatchError).
    Yields (forward text, header offsets, header lengths, fwd, rev) with fwd / rev = (bases u8, quals u8, seq_off, qual_off,
    lengths, qual_base)."""
    fastq = bool(args.forward_fastq)
    if fastq:
        yield from _read_fastq_pair_batches(args, lower_n_ambiguous)
        return
    files = [(open_input(args.forward_fasta), args.forward_fasta), (open_input(args.forward_qual), args.forward_qual),
             (open_input(args.reverse_fasta), args.reverse_fasta), (open_input(args.reverse_qual), args.reverse_qual)]
    per = 2
    carries = [b""] * len(files)
    while True:
        text0, carries[0], n_lines, eof = _whole_records(files[0][0], carries[0], per)
        texts = [text0]
        for k in range(1, len(files)):
            t, carries[k], got, _ = _whole_records(files[k][0], carries[k], per, n_lines)
            texts.append(t)
        if not text0.strip():
            if any(t.strip() for t in texts[1:]) or any(c.strip() for c in carries[1:]):
                raise NameMismatchError("", "(the forward file ended first)")
            break

        def parse(idx):
            try:
                _, qslab, off, ln, hoff, hlen, soff = parse_fasta_qual(texts[2 * idx], texts[2 * idx + 1], lower_n_ambiguous)
                return texts[2 * idx], hoff, hlen, (np.frombuffer(texts[2 * idx], dtype=np.uint8), qslab, soff, off, ln, 0)
            except MoiraError as exc:
                raise _parse_error_for(exc, files[2 * idx][1]) from None

        ftext, fh_off, fh_len, fwd = parse(0)
        rtext, rh_off, rh_len, rev = parse(1)
        n = len(fwd[4])
        if len(rev[4]) != n:
            raise NameMismatchError("(%d forward records)" % n, "(%d reverse records)" % len(rev[4]))
        _check_pair_headers(ftext, fh_off, fh_len, rtext, rh_off, rh_len)                  # moira.py:1141-1142, 1199-1200
        yield ftext, fh_off, fh_len, fwd, rev
        if eof:
            break




# ---- the reads of a run, as arrays ------------------------------------------------------------------------
REASON_OVERLAP = 4   # host-side: "overlap length below" (moira.py:886-897); the device knows reasons 0..3


class ReadSet:
    """Every read of the run: where its header / bases / qualities lie (absolute addresses into buffers kept alive in
    `keep`), and what the device said about it.  Filled batch by batch (or shard by shard), concatenated once."""

    def __init__(self, qual_sub):
        self.qual_sub = qual_sub
        self.keep = []
        self.parts = []          # dicts of arrays
        self.counters = np.zeros(L.N_COUNTERS, np.uint64)

    def add(self, **arrays):
        self.parts.append(arrays)

    def finish(self):
        keys = self.parts[0].keys() if self.parts else ()
        for k in keys:
            vals = [p[k] for p in self.parts]
            setattr(self, k, None if vals[0] is None else (vals[0] if len(vals) == 1 else np.concatenate(vals)))
        if not self.parts:
            for k, dt in (("hdr_addr", np.uint64), ("hdr_len", np.uint32), ("seq_addr", np.uint64), ("qual_addr", np.uint64),
                          ("length", np.uint32), ("ee", np.float64), ("flags", np.uint8)):
                setattr(self, k, np.zeros(0, dt))
            self.stats = self.labels = None
        self.parts = []
        self.n = int(self.length.shape[0])
        return self


def _addr(buf) -> int:
    return int(buf.ctypes.data) if isinstance(buf, np.ndarray) else int(np.frombuffer(buf, dtype=np.uint8).ctypes.data)


class Writers:
    """The output files of write_results (moira.py:323-370), opened in binary mode: the native writer hands over bytes.
    Uncompressed files are written by moira_blocks_write (parallel pwrite at offsets kept here) on a writer thread, one
    batch of blocks behind the formatter; gz files the same way through moira_blocks_write_gz (BGZF members compressed on
    all host threads: any gunzip reads them); bz2 files go through Python's compressor."""

    def __init__(self, args, output_name):
        self.native = args.output_compression in ("none", "gz")
        self.gz = args.output_compression == "gz"       # BGZF members compressed on all host threads (moira_blocks_write_gz)
        # level 1 = the library's own compressor (every piece inflated and compared before it is written): 3.6 s per 10 M reads
        # against 13.8 s for zlib's level 6, files 15 % larger; MOIRA_B200_GZ_LEVEL=2..9 selects zlib at that level
        self.gz_level = int(os.environ.get("MOIRA_B200_GZ_LEVEL", "1"))
        opener, suffix = {"none": (open, ""), "gz": (open, ".gz"), "bz2": (bz2.open, ".bz2")}[args.output_compression]
        self.files, self.names, self.by_block, self.pos = [], [], {}, {}
        self.seconds = 0.0

        def op(name, block):
            fh = opener(name + suffix, "wb")
            self.files.append(fh)
            self.names.append(name + suffix)
            self.by_block[block] = fh
            self.pos[block] = 0

        ext = "fastq" if args.output_format == "fastq" else "fasta"
        op("%s.qc.good.%s" % (output_name, ext), L.BLOCK_GOOD)
        if ext == "fasta":
            op("%s.qc.good.qual" % output_name, L.BLOCK_GOOD_QUAL)
        op("%s.qc.bad.%s" % (output_name, ext), L.BLOCK_BAD)
        if ext == "fasta":
            op("%s.qc.bad.qual" % output_name, L.BLOCK_BAD_QUAL)
        if args.collapse and args.pipeline == "mothur":
            op("%s.qc.good.names" % output_name, L.BLOCK_GOOD_NAMES)
            op("%s.qc.bad.names" % output_name, L.BLOCK_BAD_NAMES)
        if args.paired:                                                           # moira.py:366-368
            op("%s.contigs.report" % output_name, L.BLOCK_REPORT)
            head = b"header\tn_seqs\toverlap_length\tgaps\tmismatches\n"
            fh = self.by_block[L.BLOCK_REPORT]
            if self.gz:
                from .api import gz_deflate
                self.pos[L.BLOCK_REPORT] = gz_deflate(head, fh.fileno(), 0)
            else:
                fh.write(head)
                if self.native:
                    fh.flush()
                    self.pos[L.BLOCK_REPORT] = len(head)
        self._pool = ThreadPoolExecutor(1)
        self._pending = None

    def _write_now(self, blocks):
        t0 = time.time()
        try:
            for which, fh in self.by_block.items():
                if self.gz:
                    self.pos[which] += blocks.pwrite_gz(which, fh.fileno(), self.pos[which], self.gz_level)
                elif self.native:
                    self.pos[which] += blocks.pwrite(which, fh.fileno(), self.pos[which])
                else:
                    blocks.write(which, fh)
        finally:
            blocks.recycle()
            self.seconds += time.time() - t0

    def write(self, blocks):
        """Queue one batch of formatted blocks (closed when written); returns once the batch before it is on file."""
        self.drain()
        self._pending = self._pool.submit(self._write_now, blocks)

    def drain(self):
        if self._pending is not None:
            pending, self._pending = self._pending, None
            pending.result()

    def close(self):
        try:
            self.drain()
        finally:
            self._pool.shutdown(wait=True)
            L.lib.moira_blocks_free(None)       # the recycled batches
            if self.gz:                         # the empty member that ends a BGZF file (an empty output is a valid gzip file too)
                import ctypes
                n = ctypes.c_uint64()
                for which, fh in self.by_block.items():
                    L.lib.moira_gz_eof(fh.fileno(), self.pos[which], ctypes.byref(n))
            for fh in self.files:
                fh.close()


def resolve_devices(args):
    """--devices all | i,j,...  (one context and one host thread per GPU); --device i is the single-GPU spelling."""
    spec = getattr(args, "devices", None)
    if not spec:
        return [args.device]
    if spec == "all":
        import ctypes
        n = ctypes.c_int()
        L.check(L.lib.moira_device_count(ctypes.byref(n)))
        # "all" = as many of the GPUs as the input can keep busy: a context costs 0.3 s to create and tear down and the run
        # is output-bound beyond a few GB of text per GPU (measured: 10 M reads take 2.6 s on one B200, 5.1 s on eight), so one
        # GPU per 2 GiB of (inflated) input text; an explicit list is taken as given
        per_gpu = int(os.environ.get("MOIRA_B200_CLI_BYTES_PER_GPU", str(2 << 30)))
        total = 0
        for name in (args.forward_fastq, args.reverse_fastq, args.forward_fasta, args.forward_qual, args.reverse_fasta, args.reverse_qual):
            if name and os.path.exists(name):
                size = os.path.getsize(name)
                with io.open(name, "rb") as fh:
                    magic = fh.read(3)
                total += size * (4 if magic.startswith(b"\x1f\x8b\x08") or magic.startswith(b"\x42\x5a\x68") else 1)
        want = max(1, -(-total // max(1, per_gpu)))
        return list(range(max(1, min(n.value, want))))
    return [int(x) for x in spec.split(",") if x != ""]


def reduce_run_counters(ctxs, per_ctx, out):
    """The path's only exchange: per-GPU counters -> their sum.  All contexts of a CLI run live in this one
    process, so the sum is taken on the host: an NCCL communicator for 640 bytes costs seconds to build and to tear down
    (measured: 3.3 s + 3 s on two B200s against 2.5 s for the whole 10 M-read run).  Hosts with one process per GPU use
    moira_comm_init + moira_reduce_counters (NCCL over NVLink, inside the library) as bench.py does; MOIRA_B200_CLI_NCCL=1
    makes the CLI take that route too (moira_comm_init_all / moira_reduce_counters_all)."""
    if len(ctxs) == 1:
        return per_ctx[0]
    if os.environ.get("MOIRA_B200_CLI_NCCL") != "1" or len({c.device for c in ctxs}) < len(ctxs):
        return np.sum(per_ctx, axis=0).astype(np.uint64)
    from .api import comm_init_all, reduce_counters_all
    try:
        comm_init_all(ctxs)
        return reduce_counters_all(ctxs, [c.copy() for c in per_ctx])[0]
    except MoiraError as exc:
        if exc.code != L.ERR_NCCL:
            raise
        print("warning: NCCL is not available (%s); the %d counters were added on the host" % (exc.message, L.N_COUNTERS), file=out)
        return np.sum(per_ctx, axis=0).astype(np.uint64)


# ---- the three input flows --------------------------------------------------------------------------------------
def run_fastq(args, ctxs, params, lower_n, out):
    """Single-end FASTQ: the text is cut at record starts into one shard per GPU (contiguous chunks: concatenating the
    shard outputs restores input order, moira.py:455-487); every shard goes through moira_filter_fastq_ex on its own
    context and host thread -- text -> H2D -> parser, slab conversion, filter (and, with --collapse, the exact
    dereplication) on the device, three chunks in flight per GPU."""
    from .api import fastq_headers, fastq_split
    tp = time.time()
    text, keep = load_text(args.forward_fastq)
    tp = _phase("load_text", tp)
    rs = ReadSet(args.fastq_offset)
    rs.keep.append((text, keep))
    D = len(ctxs)
    cuts = fastq_split(text, D) if D > 1 and text.size else np.array([0, text.size], np.uint64)
    want_labels = bool(args.collapse)
    head = text[:1 << 20]
    lines = int(np.count_nonzero(head == 10))
    per_byte = (lines / 4.0) / max(1, head.size)

    def work(i):
        c0, c1 = int(cuts[i]), int(cuts[i + 1])
        sub = text[c0:c1]
        if sub.size == 0:
            return None
        est = int(per_byte * sub.size * 1.08) + 4096
        labels = want_labels
        for _attempt in range(3):
            try:
                return ctxs[i].filter_fastq_ex(sub, params, args.fastq_offset, max_reads=est, offsets=True, labels=labels)
            except MoiraError as exc:
                if "more than max_reads" in exc.message:
                    est = None                                     # count the records, then once more
                elif exc.code == L.ERR_NOMEM and labels:
                    labels = False                                 # sequences do not fit on the device: dereplicate on the host
                else:
                    _raise_parse_error(exc, args.forward_fastq)
        raise RuntimeError("could not size the outputs of shard %d" % i)

    if D > 1:
        with ThreadPoolExecutor(D) as pool:
            results = list(pool.map(work, range(D)))
    else:
        results = [work(0)]
    tp = _phase("moira_filter_fastq_ex (all GPUs)", tp)
    base_addr = _addr(text) if text.size else 0
    read_base = 0
    per_ctx = []
    all_labelled = want_labels
    for i, r in enumerate(results):
        if r is None:
            per_ctx.append(np.zeros(L.N_COUNTERS, np.uint64))
            continue
        c0 = np.uint64(int(cuts[i]))
        n = int(r.lengths.shape[0])
        seq_off, qual_off = r.seq_off + c0, r.qual_off + c0
        hoff, hlen = fastq_headers(text, seq_off)
        if r.labels is None:
            all_labelled = False
        rs.add(hdr_addr=hoff + np.uint64(base_addr), hdr_len=hlen, seq_addr=seq_off + np.uint64(base_addr),
               qual_addr=qual_off + np.uint64(base_addr), length=r.lengths, ee=r.filter.ee, flags=r.filter.flags, stats=None,
               labels=None if r.labels is None else r.labels + np.uint32(read_base))
        per_ctx.append(r.filter.counters)
        read_base += n
    rs.finish()
    if want_labels and not all_labelled:
        rs.labels = None
    rs.label_shards = D if (rs.labels is not None and D > 1) else 1
    tp = _phase("headers + read set", tp)
    rs.counters = reduce_run_counters(ctxs, per_ctx, out)
    tp = _phase("counters all-reduce", tp)
    return rs


def _device_round_robin(ctxs, jobs, max_in_flight=2):
    """Run job(ctx) for every job, job k on context k mod D (one host thread per GPU, at most max_in_flight jobs queued
    per GPU), yielding the results in job order: the parser of the next batches overlaps the GPUs' work."""
    D = len(ctxs)
    pools = [ThreadPoolExecutor(1) for _ in range(D)]
    pending = []
    try:
        for k, job in enumerate(jobs):
            pending.append(pools[k % D].submit(job, ctxs[k % D]))
            while len(pending) > max_in_flight * D:
                yield pending.pop(0).result()
        while pending:
            yield pending.pop(0).result()
    finally:
        for p in pools:
            p.shutdown(wait=True)


def run_fasta_qual(args, ctxs, params, lower_n, out):
    """Single-end FASTA + QUAL: whole-record blocks of both files -> native parser (all host threads) -> slab -> one
    moira_filter_batch per block, blocks dealt round-robin to the GPUs."""
    rs = ReadSet(0)
    per_ctx = [np.zeros(L.N_COUNTERS, np.uint64) for _ in ctxs]

    def jobs():
        for k, (ftext, hoff, hlen, soff, qslab, slab, offsets, lengths) in enumerate(
                read_fasta_qual_batches(args.forward_fasta, args.forward_qual, lower_n)):
            def job(ctx, k=k, ftext=ftext, hoff=hoff, hlen=hlen, soff=soff, qslab=qslab, slab=slab, offsets=offsets, lengths=lengths):
                res = ctx.filter_batch(slab, offsets, lengths, params)
                return k, ftext, hoff, hlen, soff, qslab, offsets, lengths, res
            yield job

    for k, ftext, hoff, hlen, soff, qslab, offsets, lengths, res in _device_round_robin(ctxs, jobs()):
        fa = _addr(ftext)
        rs.keep.append((ftext, qslab))
        rs.add(hdr_addr=hoff + np.uint64(fa), hdr_len=hlen, seq_addr=soff + np.uint64(fa), qual_addr=offsets + np.uint64(_addr(qslab)),
               length=lengths, ee=res.ee, flags=res.flags, stats=None, labels=None)
        per_ctx[k % len(ctxs)] += res.counters
    rs.finish()
    rs.label_shards = 1
    rs.counters = reduce_run_counters(ctxs, per_ctx, out)
    return rs


def run_pairs(args, ctxs, params, lower_n, contig_params, out):
    """Paired input: contigs and the filter on them in one call per block (process_data, moira.py:791-833), blocks dealt
    round-robin to the GPUs; --only_contig skips the filter."""
    from .api import FilterResult
    rs = ReadSet(0)
    per_ctx = [np.zeros(L.N_COUNTERS, np.uint64) for _ in ctxs]

    def jobs():
        for k, (ftext, hoff, hlen, fwd, rev) in enumerate(read_pair_batches(args, lower_n)):
            def job(ctx, k=k, ftext=ftext, hoff=hoff, hlen=hlen, fwd=fwd, rev=rev):
                pr = ctx.filter_pairs(fwd[0], fwd[1], fwd[2], fwd[4], rev[0], rev[1], rev[2], rev[4], contig_params,
                                      None if args.only_contig else params, lower_n, fwd[3], rev[3], fwd[5])
                return k, ftext, hoff, hlen, rev, pr
            yield job

    for k, ftext, hoff, hlen, rev, pr in _device_round_robin(ctxs, jobs()):
        bad = np.flatnonzero(pr.status)
        if bad.size:
            r = int(bad[0])
            st = int(pr.status[r])
            header = bytes(ftext[int(hoff[r]):int(hoff[r]) + int(hlen[r])]).decode("latin-1").replace(":", "_")
            if st == L.PAIR_BAD_BASE:                                          # moira.py:1228-1229
                seq = rev[0][int(rev[2][r]):int(rev[2][r]) + int(rev[4][r])].tobytes().decode("latin-1")
                wrong = [c for c in seq if c not in "ACTGNWSRYMKBVDH-."]
                raise ValueError('"%s" is not a recognizable IUPAC-coded base.' % (wrong[0] if wrong else "?"))
            raise ValueError("Contig construction failed for %s (%s)" % (header, {
                L.PAIR_EMPTY: "empty read", L.PAIR_BAD_QUALITY: "a quality score outside 0..252",
                L.PAIR_TOO_LONG: "reverse read longer than 1024 bases"}.get(st, "status %d" % st)))
        n = int(pr.contig_len.shape[0])
        stride = int(pr.contig_seq.shape[1]) if n else 16
        if pr.filter is None:                                                  # expected_errors = 0 (moira.py:809-810)
            res = FilterResult(np.zeros(n), np.zeros(n, np.int32), np.full(n, L.FLAG_ACCEPT, np.uint8), np.zeros(L.N_COUNTERS, np.uint64))
            if args.truncate:
                res.flags[pr.contig_len < args.truncate] = L.REASON_LENGTH << 1
        else:
            res = pr.filter
        fa = _addr(ftext)
        rows = np.arange(n, dtype=np.uint64) * np.uint64(stride)
        rs.keep.append((ftext, pr))
        rs.add(hdr_addr=hoff + np.uint64(fa), hdr_len=hlen, seq_addr=rows + np.uint64(_addr(pr.contig_seq)),
               qual_addr=rows + np.uint64(_addr(pr.contig_qual)), length=pr.contig_len, ee=res.ee, flags=res.flags,
               stats=np.stack([pr.overlap, pr.gaps, pr.mismatches], axis=1), labels=None)
        per_ctx[k % len(ctxs)] += res.counters
    rs.finish()
    rs.label_shards = 1
    rs.counters = reduce_run_counters(ctxs, per_ctx, out)
    return rs


# ---- decisions -> groups -> files (moira.py:455-519, 842-970) ----------------------------------------------------
def _string_at(addr, n):
    import ctypes
    return ctypes.string_at(int(addr), int(n)).decode("latin-1")


_T_PHASES = []


def _phase(name, t0):
    """MOIRA_B200_CLI_TIMING=1: wall-clock seconds of the host phases, printed with the summary."""
    _T_PHASES.append((name, time.time() - t0))
    return time.time()


def finish_run(args, rs, writers, out, ctx=None):
    """The collapse dictionary, the abundance sort, write_results' precedence of rules and the final counts -- on arrays;
    formatting is moira_format_records (native, all host threads).  Returns (processed, errors, minlength, minoverlap)."""
    from .api import RecordView, collapse_labels, format_records
    n = rs.n
    if n == 0:
        return 0, 0, 0, 0
    tp = time.time()
    numeric = (rs.flags & L.FLAG_NUMERIC) != 0
    if numeric.any() or np.isnan(rs.ee).any():                                   # moira.py:456-457
        bad = int(np.flatnonzero(numeric | np.isnan(rs.ee))[0])
        raise ReturnedNaNError("Error calculation failed for sequence %s" % _string_at(rs.hdr_addr[bad], rs.hdr_len[bad]).replace(":", "_"))
    accept = (rs.flags & L.FLAG_ACCEPT).astype(np.uint8)
    reason = ((rs.flags & L.FLAG_REASON_MASK) >> 1).astype(np.uint8)
    out_len = np.minimum(rs.length, np.uint32(args.truncate)) if args.truncate else rs.length           # moira.py:806-807
    # the rules between the length check and the error filter (moira.py:886-908): min_overlap (overlap_length is 0 for
    # reads that were not assembled, as in the reference), then only_contig
    not_short = reason != L.REASON_LENGTH
    if args.min_overlap:
        ov = rs.stats[:, 0] if rs.stats is not None else np.zeros(n, np.int32)
        low = not_short & (ov < args.min_overlap)
        accept[low], reason[low] = 0, REASON_OVERLAP
        not_short &= ~low
    if args.only_contig:
        accept[not_short], reason[not_short] = 1, L.REASON_NONE
    notes = {L.REASON_ERRORS: ("\terrors > %.2f" % args.maxerrors) if args.maxerrors else ("\tuncert > %.3f" % args.uncert),
             L.REASON_LENGTH: "\tlength below %s" % args.truncate, L.REASON_AMBIGS: "\tcontains ambiguities",
             REASON_OVERLAP: "\toverlap length below %s" % (args.truncate if args.output_format == "fastq" else args.min_overlap)}
    rec = RecordView(None, rs.hdr_addr, rs.hdr_len, None, rs.seq_addr, None, rs.qual_addr, out_len, rs.qual_sub)
    stats = None if rs.stats is None else (rs.stats[:, 0], rs.stats[:, 1], rs.stats[:, 2])
    common = dict(fastq=args.output_format == "fastq", fastq_offset=args.fastq_offset, usearch=args.pipeline == "USEARCH",
                  relabel=args.relabel, notes=notes, stats=stats)
    CHUNK = 1 << 18
    if args.collapse:
        # moira.py:459-475 + 491-504: groups by first appearance, representative = first read with the strictly smallest
        # ee, names in the reference's order, output by abundance
        labels = rs.labels
        if labels is not None and rs.label_shards > 1:
            # every GPU dereplicated its own shard exactly; sequences that span shards meet here: only the shards' owner
            # reads (one per distinct sequence and shard) are compared on the host
            owners = np.flatnonzero(labels == np.arange(n, dtype=np.uint32))
            oc = collapse(None, rs.seq_addr[owners], out_len[owners], np.zeros(owners.size))
            lut = np.arange(n, dtype=np.uint32)
            lut[owners] = owners[oc.rep[oc.group_of_read]].astype(np.uint32)
            labels = lut[labels]
        tp = _phase("decision arrays", tp)
        if labels is None and ctx is not None and n < (1 << 31):
            # no labels from the input flow (pairs, FASTA + QUAL, or a shard that did not fit): make them on the GPU from
            # wherever the sequences lie in host memory
            try:
                labels = ctx.collapse_addr(rs.seq_addr, out_len)
            except MoiraError as exc:
                if exc.code != L.ERR_NOMEM:
                    raise
        if labels is None:
            col = collapse(None, rs.seq_addr, out_len, rs.ee)
        else:
            col = None
            if ctx is not None and n < (1 << 31):
                try:
                    v = ctx.collapse_groups(labels, rs.ee)         # groups / representatives / orders on the GPU (uint32 views)
                    from .api import CollapseResult
                    col = CollapseResult(*[np.asarray(getattr(v, f), dtype=np.uint64) for f in
                                           ("group_of_read", "rep", "size", "member_start", "members", "order")])
                except MoiraError as exc:
                    if exc.code != L.ERR_NOMEM:
                        raise
            if col is None:
                col = collapse_labels(labels, rs.ee)
        sel, sel_group = col.rep[col.order], col.order
        sizes = col.size[col.order].astype(np.int64)
        tp = _phase("groups", tp)
        for k0 in range(0, sel.size, CHUNK):
            k1 = min(sel.size, k0 + CHUNK)
            blocks = format_records(rec, sel[k0:k1], rs.ee, accept, reason, names=args.pipeline == "mothur", first_index=1 + k0,
                                    sel_group=sel_group[k0:k1], member_start=col.member_start, members=col.members, **common)
            writers.write(blocks)
        acc_s, rsn_s = accept[sel], reason[sel]
    else:
        tp = _phase("decision arrays", tp)
        for k0 in range(0, n, CHUNK):
            k1 = min(n, k0 + CHUNK)
            blocks = format_records(rec, np.arange(k0, k1, dtype=np.uint64), rs.ee, accept, reason, first_index=k0, **common)
            writers.write(blocks)
        acc_s, rsn_s, sizes = accept, reason, np.ones(n, np.int64)
    writers.drain()
    tp = _phase("format + write (writer thread busy %.2f s)" % writers.seconds, tp)
    tp = time.time()
    rej = acc_s == 0
    minlength = int(sizes[rej & (rsn_s == L.REASON_LENGTH)].sum())
    minoverlap = int(sizes[rej & (rsn_s == REASON_OVERLAP)].sum())
    errors = int(sizes[rej].sum()) - minlength - minoverlap
    _phase("final counts", tp)
    return n, errors, minlength, minoverlap


# ---- main (moira.py:264-578) ---------------------------------------------------------------------------
def main(args, out=sys.stdout) -> int:
    if not args.silent:
        print("\nmoira_b200 %s -- Poisson-binomial read filtering on NVIDIA B200\n" % __version__, file=out)
    if not check_arguments(args, out):
        return 1
    if args.output_prefix:
        output_name = args.output_prefix
    elif args.forward_fastq:
        output_name = ".".join(args.forward_fastq.split(".")[:-1])
    else:
        output_name = ".".join(args.forward_fasta.split(".")[:-1])

    calc = "poisson_binomial" if args.error_calc == "poisson_binomial_py" else args.error_calc
    params = FilterParams(error_calc=calc, alpha=args.alpha, uncert=args.uncert, maxerrors=args.maxerrors,
                          ambigs=args.ambigs, round=args.round, truncate=args.truncate,
                          exact_ee=bool(args.collapse) or args.pipeline == "USEARCH", ee_output="final")
    lower_n = args.error_calc == "poisson_binomial"      # bernoullimodule.c:196 vs moira.py:1605/1660
    contig_params = ContigParams(match=args.match, mismatch=args.mismatch, gap=args.gap, insert=args.insert, deltaq=args.deltaq,
                                 consensus_qscore=args.consensus_qscore, qscore_cap=args.qscore_cap, trim_overlap=args.trim_overlap)
    try:
        inputs = [args.forward_fastq] if args.forward_fastq else [args.forward_fasta, args.forward_qual]
        if args.paired:
            inputs += [args.reverse_fastq] if args.forward_fastq else [args.reverse_fasta, args.reverse_qual]
        for name in inputs:
            io.open(name, "rb").close()
        writers = Writers(args, output_name)
    except IOError as exc:
        print(exc, file=out)
        return 1

    t0 = time.time()
    del _T_PHASES[:]
    ctxs = []
    try:
        devs = resolve_devices(args)
        if len(devs) > 1:
            # a CUDA primary context per GPU costs 0.3-2 s to create: side by side, not one after the other
            with ThreadPoolExecutor(len(devs)) as pool:
                futs = [pool.submit(Context, d) for d in devs]
                for f in futs:
                    try:
                        ctxs.append(f.result())
                    except Exception:
                        for g in futs:
                            if g is not f and g.exception() is None:
                                g.result().close()
                        for c in ctxs:
                            c.close()
                        ctxs = []
                        raise
        else:
            ctxs = [Context(devs[0])]
        _phase("contexts", t0)
        if args.paired:
            rs = run_pairs(args, ctxs, params, lower_n, contig_params, out)
        elif args.forward_fastq:
            rs = run_fastq(args, ctxs, params, lower_n, out)
        else:
            rs = run_fasta_qual(args, ctxs, params, lower_n, out)
        t_filter = time.time() - t0
        processed, discarded_errors, discarded_minlength, discarded_minoverlap = finish_run(args, rs, writers, out, ctxs[0] if ctxs else None)
    finally:
        tc = time.time()
        writers.close()
        tc = _phase("close writers", tc)
        for c in ctxs:
            c.close()
        _phase("close contexts", tc)

    if os.environ.get("MOIRA_B200_CLI_TIMING") == "1":
        for name, sec in _T_PHASES:
            print("  [timing] %-60s %.3f s" % (name, sec), file=out)
    if not args.silent and processed:
        print("%d sequences processed in %.1f seconds (%.1f s to the decisions, %d GPU%s)." %
              (processed, time.time() - t0, t_filter, len(ctxs), "" if len(ctxs) == 1 else "s"), file=out)
        remaining = processed - discarded_errors - discarded_minlength - discarded_minoverlap
        print("\n- Kept %d (%.2f%%) of the original sequences." % (remaining, remaining / processed * 100), file=out)
        if args.truncate:
            print("- %d (%.2f%%) of the original sequences were discarded due to length < %s." %
                  (discarded_minlength, discarded_minlength / processed * 100, args.truncate), file=out)
        if args.min_overlap and args.paired:
            print("- %d (%.2f%%) of the original sequences were discarded due to overlap length < %s." %
                  (discarded_minoverlap, discarded_minoverlap / processed * 100, args.min_overlap), file=out)
        print("- %d (%.2f%%) of the original sequences were discarded due to low quality.\n" %
              (discarded_errors, discarded_errors / processed * 100), file=out)
        if not args.collapse:
            c = rs.counters
            print("- device counters, summed over %d GPU%s: %d reads, %d accepted, %d rejected for errors, %d for length, %d for "
                  "ambiguities, %d within 1e-12 of the cutoff." % (len(ctxs), "" if len(ctxs) == 1 else "s", int(c[L.CNT_READS]),
                  int(c[L.CNT_ACCEPTED]), int(c[L.CNT_BAD_ERRORS]), int(c[L.CNT_BAD_LENGTH]), int(c[L.CNT_BAD_AMBIGS]),
                  int(c[L.CNT_NEAR_CUTOFF])), file=out)
        print("The following output files were generated:", file=out)
        for name in writers.names:
            print(name, file=out)
    return 0


def run(argv=None) -> int:
    return main(parse_arguments(argv))


if __name__ == "__main__":
    sys.exit(run())
