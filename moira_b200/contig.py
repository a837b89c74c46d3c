"""Paired-end contig constructor with the reference's function names and error behaviour, computed by the
CUDA path:

    reverse_complement(sequence, quals=None)                      moira/moira.py:1207-1236  (host, string work only)
    nw_align(seq_1, seq_2, match, mismatch, gap)                  moira/nw_align.pyx:49     (moira_nw_align)
    make_contig(forward_aligned, forward_quals, reverse_aligned,
                reverse_quals, insert, deltaq, consensus_qscore,
                qscore_cap, trim_overlap)                         moira/moira.py:1375       (moira_make_contig)

One call = one kernel launch on a batch of one; batches go through Context.filter_pairs.
There is no CPU fallback: without a usable GPU the first call raises.
"""
from __future__ import annotations

from . import _lib as L
from .api import ContigParams, Context, MoiraError

_COMPLEMENT = {"A": "T", "C": "G", "T": "A", "G": "C", "N": "N", "W": "W", "S": "S", "R": "Y", "Y": "R", "M": "K",
               "K": "M", "B": "V", "V": "B", "D": "H", "H": "D", "-": "-", ".": "."}          # moira.py:1210-1213
_ctx = None


class LengthMismatchError(Exception):
    """moira.py:1024-1038: raised bare by the calculators and the contig constructor, with a header and file names by the parsers."""

    def __init__(self, header=None, *files):
        self.header, self.files = header, files
        if header is None:
            super().__init__("Sequence and qualities are of different lengths.")
        else:
            super().__init__("Sequence and quality lengths differ for %s (%s)" % (header, ", ".join(map(str, files))))


def _context() -> Context:
    global _ctx
    if _ctx is None:
        _ctx = Context(0)
    return _ctx


def reverse_complement(sequence, quals=None):
    """Returns the reverse complement of a sequence (and the reversed qualities)."""
    sequence = str(sequence)
    if quals:
        quals = list(quals)
        for x in quals:
            if x not in (".", "-"):
                int(x)
        if len(sequence.replace("-", "").replace(".", "")) != len([x for x in quals if x not in (".", "-")]):
            raise LengthMismatchError                                                            # moira.py:1219-1220
    out = []
    for base in sequence[::-1]:
        try:
            out.append(_COMPLEMENT[base])
        except KeyError:
            raise ValueError('"%s" is not a recognizable IUPAC-coded base.' % base) from None   # moira.py:1229
    if quals:
        quals.reverse()
        return "".join(out), quals
    return "".join(out)


def nw_align(seq_1, seq_2, match, mismatch, gap, refine_overlap=True, verbose=False):
    """Needleman-Wunsch with mothur's overlap refinement -> (seq_1_aligned, seq_2_aligned, score)."""
    if not refine_overlap:
        raise NotImplementedError("only refine_overlap=True is built: it is the only mode moira itself uses (moira.py:794-798)")
    return _context().nw_align(str(seq_1), str(seq_2), int(match), int(mismatch), int(gap))


def make_contig(forward_aligned, forward_quals, reverse_aligned, reverse_quals, insert, deltaq, consensus_qscore, qscore_cap,
                trim_overlap):
    """-> (contig, contig_quals, overlap_length, gaps, mismatches)."""
    forward_quals = [int(q) for q in forward_quals]
    reverse_quals = [int(q) for q in reverse_quals]
    params = ContigParams(insert=int(insert), deltaq=int(deltaq), consensus_qscore=consensus_qscore, qscore_cap=qscore_cap,
                          trim_overlap=bool(trim_overlap))
    try:
        return _context().make_contig(str(forward_aligned), forward_quals, str(reverse_aligned), reverse_quals, params)
    except MoiraError as exc:
        if exc.code == L.ERR_LENGTH_MISMATCH:
            raise LengthMismatchError from None                                                  # moira.py:1407-1410
        if exc.code in (L.ERR_BAD_ARG, L.ERR_BAD_QUALITY):
            raise ValueError(exc.message) from None                                              # moira.py:1405-1416
        raise
