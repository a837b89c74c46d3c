"""moira_b200 command line: the Python-3 host of the B200 path.

Keeps the reference CLI's flags, defaults, input sniffing, collapse logic, output files and
formatting (fpusan/moira v1.3.2, moira/moira.py: parse_arguments :581-675, check_arguments :678-781,
main :264-578, write_results :842-970, parsers :1058-1204), but instead of calling
`bernoulli.calculate_errors_PB` once per read (moira.py:817) it parses reads in batches straight into
quality slabs and runs them through libmoira_b200.so (one moira_filter_batch per batch).

Scope: single-end filtering (`--forward_fastq` or `--forward_fasta` + `--forward_qual`).  Paired-end
contig construction (`--paired`, `--only_contig`) is upstream of the hot path and not part of this
build; those flags are accepted and rejected with a message.
`--error_calc` adds `expected_error` (north_star) and drops `bootstrap` (deprecated, moira.py:772-776);
`poisson_binomial_py` is an alias of `poisson_binomial` with the Python twin's 'N'-only rule.
"""
from __future__ import annotations

import argparse
import bz2
import gzip
import io
import sys
import time

import numpy as np

from . import _lib as L
from .api import ContigParams, Context, FilterParams, MoiraError, collapse, pack_reads, parse_fasta_qual, parse_fastq

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


class LengthMismatchError(Exception):
    def __init__(self, header, *files):
        super().__init__("Sequence and quality lengths differ for %s (%s)" % (header, ", ".join(map(str, files))))


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
                         help="Accepted for compatibility; the GPU path does not use worker processes.")
    general.add_argument("--paired", action="store_true")
    general.add_argument("-fo", "--fastq_offset", type=int, default=33)
    general.add_argument("--only_contig", action="store_true")
    general.add_argument("--silent", action="store_true")
    general.add_argument("--nowarnings", action="store_true")
    general.add_argument("--doc", action="store_true")
    general.add_argument("--device", type=int, default=0, help="CUDA device index (B200 path only).")
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


def read_fastq_batches(fh, fastq_offset, lower_n_ambiguous, fname):
    """Yield (headers, seqs, quals_arrays, slab, offsets, lengths) per batch of whole 4-line records."""
    carry = b""
    while True:
        block = fh.read(BATCH_BYTES)
        data = carry + block
        if not data:
            break
        if block:
            # keep whole records: cut at the last newline that ends a multiple of 4 lines
            n_lines = data.count(b"\n")
            keep_lines = n_lines - (n_lines % 4)
            if keep_lines == 0:
                carry = data
                continue
            pos = -1
            # find position after keep_lines-th newline
            idx = np.flatnonzero(np.frombuffer(data, dtype=np.uint8) == 10)
            pos = int(idx[keep_lines - 1]) + 1
            text, carry = data[:pos], data[pos:]
        else:
            text, carry = data, b""
        try:
            slab, offsets, lengths, hoff, hlen, soff, qoff = parse_fastq(text, fastq_offset, lower_n_ambiguous)
        except MoiraError as exc:
            if exc.code == L.ERR_PARSE:
                name = exc.message.split(":")[0]
                cls = {"EmptySeqError": EmptySeqError, "EmptyQualError": EmptyQualError}.get(name)
                if cls:
                    raise cls(exc.message, fname) from None
                raise LengthMismatchError(exc.message, fname) from None
            raise
        n = len(lengths)
        headers = [text[int(hoff[i]):int(hoff[i]) + int(hlen[i])].decode("latin-1").replace(":", "_") for i in range(n)]
        seqs = [text[int(soff[i]):int(soff[i]) + int(lengths[i])].decode("latin-1") for i in range(n)]
        # qualities for the writers, as process_data returns them (Q <= 0 -> 1, moira.py:814)
        buf = np.frombuffer(text, dtype=np.uint8)
        quals = []
        for i in range(n):
            q = buf[int(qoff[i]):int(qoff[i]) + int(lengths[i])].astype(np.int32) - fastq_offset
            quals.append(np.where(q > 0, q, 1))
        yield headers, seqs, quals, slab, offsets, lengths
        if not block:
            break


def _cut_lines(data: bytes, n_lines: int) -> int:
    """Byte position just after the n_lines-th newline of data (len(data) if it has fewer)."""
    idx = np.flatnonzero(np.frombuffer(data, dtype=np.uint8) == 10)
    return int(idx[n_lines - 1]) + 1 if 0 < n_lines <= len(idx) else len(data)


def read_fasta_qual_batches(ffh, qfh, lower_n_ambiguous, fasta_name, qual_name):
    """moira.py:1093-1149, single-end: sequences and qualities on one line each.  Whole-record blocks of
    both files go through the native parser (moira_parse_fasta_qual)."""
    fcarry = qcarry = b""
    while True:
        block = ffh.read(BATCH_BYTES)
        fdata = fcarry + block
        if block:
            n_lines = fdata.count(b"\n")
            keep = n_lines - (n_lines % 2)
            if keep == 0:
                fcarry = fdata
                continue
            pos = _cut_lines(fdata, keep)
            ftext, fcarry = fdata[:pos], fdata[pos:]
        else:
            ftext, fcarry = fdata, b""
            keep = ftext.count(b"\n") + (1 if ftext and not ftext.endswith(b"\n") else 0)
            keep -= keep % 2
        # the same number of lines from the qual file
        while qcarry.count(b"\n") < keep:
            qblock = qfh.read(BATCH_BYTES)
            if not qblock:
                break
            qcarry += qblock
        if block:
            qpos = _cut_lines(qcarry, keep)
            qtext, qcarry = qcarry[:qpos], qcarry[qpos:]
        else:
            qtext, qcarry = qcarry + qfh.read(), b""
            if len(qtext.split()) and not ftext.strip():
                raise NameMismatchError("", _norm_header(qtext.decode("latin-1").splitlines()[0], ">"))
        if not ftext.strip():
            break
        try:
            slab, qslab, offsets, lengths, hoff, hlen, soff = parse_fasta_qual(ftext, qtext, lower_n_ambiguous)
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
        n = len(lengths)
        headers = [ftext[int(hoff[i]):int(hoff[i]) + int(hlen[i])].decode("latin-1").replace(":", "_") for i in range(n)]
        seqs = [ftext[int(soff[i]):int(soff[i]) + int(lengths[i])].decode("latin-1") for i in range(n)]
        quals = [np.maximum(qslab[int(offsets[i]):int(offsets[i]) + int(lengths[i])].astype(np.int32), 1) for i in range(n)]   # moira.py:814
        yield headers, seqs, quals, slab, offsets, lengths
        if not block:
            break


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


def read_pair_batches(args, lower_n_ambiguous):
    """Paired input (moira.py:1093-1204 with both files): blocks of whole records of the forward file and the same
    number of records of the reverse file, parsed natively, headers compared (NameMismatchError).
    Yields (headers, fwd, rev) with fwd / rev = (bases u8, quals u8, seq_off, qual_off, lengths, qual_base)."""
    fastq = bool(args.forward_fastq)
    if fastq:
        files = [(open_input(args.forward_fastq), args.forward_fastq), (open_input(args.reverse_fastq), args.reverse_fastq)]
        per = 4
    else:
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
                if fastq:
                    _, _, ln, hoff, hlen, soff, qoff = parse_fastq(texts[idx], args.fastq_offset, lower_n_ambiguous)
                    buf = np.frombuffer(texts[idx], dtype=np.uint8)
                    return texts[idx], hoff, hlen, (buf, buf, soff, qoff, ln, args.fastq_offset)
                _, qslab, off, ln, hoff, hlen, soff = parse_fasta_qual(texts[2 * idx], texts[2 * idx + 1], lower_n_ambiguous)
                return texts[2 * idx], hoff, hlen, (np.frombuffer(texts[2 * idx], dtype=np.uint8), qslab, soff, off, ln, 0)
            except MoiraError as exc:
                if exc.code == L.ERR_PARSE:
                    name = exc.message.split(":")[0]
                    cls = {"EmptySeqError": EmptySeqError, "EmptyQualError": EmptyQualError}.get(name)
                    fname = files[idx if fastq else 2 * idx][1]
                    if cls:
                        raise cls(exc.message, fname) from None
                    if name == "NameMismatchError":
                        raise NameMismatchError(exc.message, "") from None
                    raise LengthMismatchError(exc.message, fname) from None
                raise

        ftext, fh_off, fh_len, fwd = parse(0)
        rtext, rh_off, rh_len, rev = parse(1)
        n = len(fwd[4])
        if len(rev[4]) != n:
            raise NameMismatchError("(%d forward records)" % n, "(%d reverse records)" % len(rev[4]))
        headers = [ftext[int(fh_off[i]):int(fh_off[i]) + int(fh_len[i])].decode("latin-1").replace(":", "_") for i in range(n)]
        for i in range(n):                                                   # moira.py:1141-1142, 1199-1200
            rh = rtext[int(rh_off[i]):int(rh_off[i]) + int(rh_len[i])].decode("latin-1").replace(":", "_")
            if rh != headers[i]:
                raise NameMismatchError(headers[i], None, rh, None)
        yield headers, fwd, rev
        if eof:
            break


# ---- output (moira.py:842-970) ----------------------------------------------------------------------
class Writers:
    def __init__(self, args, output_name):
        opener, suffix = {"none": (open, ""), "gz": (gzip.open, ".gz"), "bz2": (bz2.open, ".bz2")}[args.output_compression]
        self.files = []

        def op(name):
            fh = opener(name + suffix, "wt")
            self.files.append(fh)
            self.names.append(name + suffix)
            return fh

        self.names = []
        if args.output_format == "fastq":
            self.good = op("%s.qc.good.fastq" % output_name)
            self.good_qual = None
            self.bad = op("%s.qc.bad.fastq" % output_name)
            self.bad_qual = None
        else:
            self.good = op("%s.qc.good.fasta" % output_name)
            self.good_qual = op("%s.qc.good.qual" % output_name)
            self.bad = op("%s.qc.bad.fasta" % output_name)
            self.bad_qual = op("%s.qc.bad.qual" % output_name)
        if args.collapse and args.pipeline == "mothur":
            self.good_names = op("%s.qc.good.names" % output_name)
            self.bad_names = op("%s.qc.bad.names" % output_name)
        else:
            self.good_names = self.bad_names = None
        if args.paired:                                                           # moira.py:366-368
            self.report = op("%s.contigs.report" % output_name)
            self.report.write("header\tn_seqs\toverlap_length\tgaps\tmismatches\n")
        else:
            self.report = None

    def close(self):
        for fh in self.files:
            fh.close()


REASON_OVERLAP = 100   # host-side: "overlap length below" (moira.py:886-897); the device knows reasons 0..3


def write_result(index, header, sequence, quals, expected_errors, names_info, accept, reason, args, w: Writers, contig_stats=None):
    """One record, formatted as write_results does (moira.py:842-970); the accept/reason pair comes
    from the device.  Returns (discarded_errors, discarded_minlength, discarded_minoverlap)."""
    if args.relabel:
        header = "%s%d" % (args.relabel, index)
    if args.pipeline == "USEARCH":
        header = header + ";ee=%.2f;size=%d;" % (expected_errors, len(names_info) if names_info else 1)
    n_members = len(names_info) if names_info else 1
    if w.report is not None and contig_stats is not None:                        # moira.py:866-870
        w.report.write("%s\t%s\t%s\t%s\t%s\n" % ((header, n_members) + tuple(contig_stats)))
        # the rules that sit between the length check and the error filter when contigs were built (moira.py:886-908)
        if reason != L.REASON_LENGTH:
            if args.min_overlap and contig_stats[0] < args.min_overlap:
                accept, reason = False, REASON_OVERLAP
            elif args.only_contig:
                accept, reason = True, L.REASON_NONE
    if accept:
        out, out_q, out_n, note = w.good, w.good_qual, w.good_names, ""
    else:
        out, out_q, out_n = w.bad, w.bad_qual, w.bad_names
        if reason == L.REASON_LENGTH:
            note = "\tlength below %s" % args.truncate
        elif reason == REASON_OVERLAP:                                            # the fastq branch prints --truncate (moira.py:888)
            note = "\toverlap length below %s" % (args.truncate if args.output_format == "fastq" else args.min_overlap)
        elif reason == L.REASON_AMBIGS:
            note = "\tcontains ambiguities"
        elif args.maxerrors:
            note = "\terrors > %.2f" % args.maxerrors
        else:
            note = "\tuncert > %.3f" % args.uncert
    if args.output_format == "fastq":
        out.write("@%s%s\n%s\n+\n%s\n" % (header, note, sequence, "".join(chr(int(q) + args.fastq_offset) for q in quals)))
    else:
        out.write(">%s%s\n%s\n" % (header, note, sequence))
        out_q.write(">%s%s\n%s\n" % (header, note, " ".join(str(int(q)) for q in quals)))
    if args.collapse and args.pipeline == "mothur" and out_n is not None:
        out_n.write("%s\t%s\n" % (header, ",".join(names_info)))
    if accept:
        return 0, 0, 0
    if reason == L.REASON_LENGTH:
        return 0, n_members, 0
    return (0, 0, n_members) if reason == REASON_OVERLAP else (n_members, 0, 0)


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
        if args.paired:
            batches = read_pair_batches(args, lower_n)
        elif args.forward_fastq:
            batches = read_fastq_batches(open_input(args.forward_fastq), args.fastq_offset, lower_n, args.forward_fastq)
        else:
            batches = read_fasta_qual_batches(open_input(args.forward_fasta), open_input(args.forward_qual), lower_n,
                                              args.forward_fasta, args.forward_qual)
        writers = Writers(args, output_name)
    except IOError as exc:
        print(exc, file=out)
        return 1

    ctx = Context(args.device)

    def results():
        """(headers, sequences, qualities, ee, accept, reason, contig statistics or None) per batch."""
        if not args.paired:
            for headers, seqs, quals, slab, offsets, lengths in batches:
                res = ctx.filter_batch(slab, offsets, lengths, params)
                yield headers, seqs, quals, res, None
            return
        for headers, fwd, rev in batches:
            # contigs and the filter on them in one call (process_data, moira.py:791-833); --only_contig skips the filter
            pr = ctx.filter_pairs(fwd[0], fwd[1], fwd[2], fwd[4], rev[0], rev[1], rev[2], rev[4], contig_params,
                                  None if args.only_contig else params, lower_n, fwd[3], rev[3], fwd[5])
            bad = np.flatnonzero(pr.status)
            if bad.size:
                r = int(bad[0])
                st = int(pr.status[r])
                if st == L.PAIR_BAD_BASE:                                          # moira.py:1228-1229
                    seq = rev[0][int(rev[2][r]):int(rev[2][r]) + int(rev[4][r])].tobytes().decode("latin-1")
                    wrong = [c for c in seq if c not in "ACTGNWSRYMKBVDH-."]
                    raise ValueError('"%s" is not a recognizable IUPAC-coded base.' % (wrong[0] if wrong else "?"))
                raise ValueError("Contig construction failed for %s (%s)" % (headers[r], {
                    L.PAIR_EMPTY: "empty read", L.PAIR_BAD_QUALITY: "a quality score outside 0..252",
                    L.PAIR_TOO_LONG: "reverse read longer than 1024 bases"}.get(st, "status %d" % st)))
            seqs, quals = [], []
            for r in range(len(headers)):
                c, q = pr.contig(r)
                seqs.append(c)
                quals.append(np.maximum(np.asarray(q, dtype=np.int32), 1))         # moira.py:814
            if pr.filter is None:                                                  # expected_errors = 0 (moira.py:809-810)
                n = len(headers)
                from .api import FilterResult
                res = FilterResult(np.zeros(n), np.zeros(n, np.int32), np.full(n, L.FLAG_ACCEPT, np.uint8), np.zeros(L.N_COUNTERS, np.uint64))
                if args.truncate:
                    short = pr.contig_len < args.truncate
                    res.flags[short] = L.REASON_LENGTH << 1
            else:
                res = pr.filter
            yield headers, seqs, quals, res, np.stack([pr.overlap, pr.gaps, pr.mismatches], axis=1)

    processed = 0
    discarded_errors = discarded_minlength = discarded_minoverlap = 0
    all_headers, all_seqs, all_quals, all_ee, all_acc, all_rsn, all_stats = [], [], [], [], [], [], []   # --collapse: kept for the epilogue
    t0 = time.time()
    try:
        for headers, seqs, quals, res, stats in results():
            if np.isnan(res.ee).any() or res.numeric.any():
                bad = int(np.flatnonzero(np.isnan(res.ee) | res.numeric)[0])
                raise ReturnedNaNError("Error calculation failed for sequence %s" % headers[bad])
            accept = res.accept
            reason = res.reason
            if args.truncate:                                                      # moira.py:806-807
                seqs = [s_[:args.truncate] for s_ in seqs]
                quals = [q_[:args.truncate] for q_ in quals]
            if args.collapse:
                all_headers += headers
                all_seqs += seqs
                all_quals += quals
                all_ee.append(res.ee.copy())
                all_acc.append(accept.copy())
                all_rsn.append(reason.copy())
                if stats is not None:
                    all_stats.append(stats)
                processed += len(headers)
            else:
                for i, header in enumerate(headers):
                    de, dl, do = write_result(processed, header, seqs[i], quals[i], float(res.ee[i]), None, bool(accept[i]),
                                              int(reason[i]), args, writers, None if stats is None else stats[i].tolist())
                    discarded_errors += de
                    discarded_minlength += dl
                    discarded_minoverlap += do
                    processed += 1
            if not args.silent:
                print("%d sequences processed in %.1f seconds.\r" % (processed, time.time() - t0), end="", file=out)
        if args.collapse and processed:
            # moira.py:459-475 + 491-504, natively: groups by first appearance, representative = first read
            # with the strictly smallest ee, names in the reference's order, output by abundance
            ee_all = np.concatenate(all_ee)
            acc_all = np.concatenate(all_acc)
            rsn_all = np.concatenate(all_rsn)
            stats_all = np.concatenate(all_stats) if all_stats else None
            seq_len = np.fromiter((len(s_) for s_ in all_seqs), dtype=np.uint32, count=processed)
            seq_off = np.zeros(processed, dtype=np.uint64)
            seq_off[1:] = np.cumsum(seq_len[:-1], dtype=np.uint64)
            col = collapse("".join(all_seqs).encode("latin-1"), seq_off, seq_len, ee_all)
            for index, g in enumerate(col.order.tolist(), start=1):
                rep = int(col.rep[g])
                names = [all_headers[r] for r in col.members[int(col.member_start[g]):int(col.member_start[g + 1])].tolist()]
                de, dl, do = write_result(index, all_headers[rep], all_seqs[rep], all_quals[rep], float(ee_all[rep]), names,
                                          bool(acc_all[rep]), int(rsn_all[rep]), args, writers,
                                          None if stats_all is None else stats_all[rep].tolist())
                discarded_errors += de
                discarded_minlength += dl
                discarded_minoverlap += do
    finally:
        writers.close()
        ctx.close()

    if not args.silent and processed:
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
        print("The following output files were generated:", file=out)
        for name in writers.names:
            print(name, file=out)
    return 0


def run(argv=None) -> int:
    return main(parse_arguments(argv))


if __name__ == "__main__":
    sys.exit(run())
