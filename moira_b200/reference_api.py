"""The functions a program (or test) that does `import moira` finds in the reference module (moira/moira.py), computed on the
CUDA path -- same names, argument order, return values and exceptions, so that the reference's own test file
(moira/test/test_moira.py) reads the same against this module (tests/test_moira_module.py is that file, test by test):

    calculate_errors_PB(sequence, quals, alpha)          moira.py:1561-1634   Python twin: only 'N' is ambiguous (:1605)
    calculate_errors_poisson(sequence, quals, alpha)     moira.py:1637-1679
    interpolate(errors1, prob1, errors2, prob2, alpha)   moira.py:1723-1737   (three operations on the host: not a kernel)
    reverse_complement / nw_align / make_contig          moira.py:1207-1558   (moira_b200.contig)
    process_data(header, fseq, fquals, rseq, rquals, args)   moira.py:784-840
    parse_arguments / check_arguments / main / open_input, the exception classes   (moira_b200.cli)

Single calls are batches of one (a kernel launch and two PCIe round trips each): this module is for parity and for small
scripts; throughput comes from the batched interface (moira_b200.api) and the command line (moira_b200.cli).
There is no CPU fallback: without a usable GPU the first computing call raises.
"""
from __future__ import annotations

import math

from . import _lib as L
from .api import Context, FilterParams, MoiraError, pack_reads
from .cli import (EmptyQualError, EmptySeqError, NameMismatchError, ReturnedNaNError, UnpairedFilesError, check_arguments,  # noqa: F401
                  open_input, parse_arguments)
from .cli import main as _cli_main
from .contig import LengthMismatchError, make_contig, nw_align, reverse_complement  # noqa: F401

_ctx = None


def _context() -> Context:
    global _ctx
    if _ctx is None:
        _ctx = Context(0)
    return _ctx


def main(args, out=None):
    """moira.py:264-578 with an argument object of the caller's own making (the reference's tests build one by hand,
    test_moira.py:33-36, 130-137): attributes it does not carry take the parser's defaults."""
    import sys
    for key, value in vars(parse_arguments([])).items():
        if not hasattr(args, key):
            setattr(args, key, value)
    return _cli_main(args, out if out is not None else sys.stdout)


def _check(sequence, quals, alpha):
    """The parameter checks both error calculators start with (moira.py:1587-1595, :1644-1652)."""
    sequence = str(sequence)
    quals = [int(q) for q in list(quals)]
    alpha = float(alpha)
    if len(sequence) != len(quals):
        raise LengthMismatchError()
    if alpha <= 0 or alpha > 1:
        raise ValueError("Alpha must be between 0 (not included) and 1.")
    return sequence, quals, alpha


def _one_read(sequence, quals, alpha, error_calc):
    if alpha >= 1:
        return 0.0, sequence.count("N")          # 1 - alpha <= 0: the first term already exceeds it, j* = 0 (moira.py:1620-1632)
    slab, off, ln = pack_reads([sequence], [quals], lower_n_ambiguous=False)           # the Python twins count 'N' only
    p = FilterParams(error_calc=error_calc, alpha=alpha, exact_ee=True, ee_output="raw", ambigs="ignore", uncert=1.0)
    res = _context().filter_batch(slab, off, ln, p)
    if int(res.flags[0]) & L.FLAG_NUMERIC:
        raise OverflowError("the statistic of this read cannot be computed (the reference's loop overflows here)")
    return float(res.ee[0]), int(res.ns[0])


def calculate_errors_PB(sequence, quals, alpha):
    """Calculate the errors on a sequence using the Poisson binomial method (sum of bernoulli random variables).
    Expects quals to be a list of integers containing Phred quality scores."""
    sequence, quals, alpha = _check(sequence, quals, alpha)
    for q in quals:
        if q < 0:
            raise ValueError("Qualities must have positive values.")                    # moira.py:1603
    try:
        return _one_read(sequence, [q if q > 0 else 1 for q in quals], alpha, "poisson_binomial")
    except MoiraError as exc:
        if exc.code == L.ERR_BAD_QUALITY:
            raise ValueError(exc.message) from None
        raise


def calculate_errors_poisson(sequence, quals, alpha):
    """Calculate the errors in a sequence approximating the sum of bernouilli random variables to a poisson distribution.
    Expects quals to be a list of integers containing Phred quality scores."""
    sequence, quals, alpha = _check(sequence, quals, alpha)
    for q in quals:
        if q < 0:
            raise ValueError("Qualities must have positive values.")                    # moira.py:1658
    try:
        return _one_read(sequence, quals, alpha, "poisson")
    except MoiraError as exc:
        if exc.code == L.ERR_BAD_QUALITY:
            raise ValueError(exc.message) from None
        raise


def interpolate(errors1, prob1, errors2, prob2, alpha):
    """Perform a linear interpolation in the errors distribution to return the number of errors that has an accumulated
    probability of 1 - alpha (moira.py:1723-1737)."""
    result = errors1 + ((errors2 - errors1) * ((1 - alpha) - prob1) / (prob2 - prob1))
    if result < 0:
        result = 0
    return result


def process_data(header, forward_sequence, forward_quals, reverse_sequence, reverse_quals, args):
    """Assemble (if needed), truncate (if needed) and quality check the sequences (moira.py:784-840):
    -> (header, contig, contig_quals, expected_errors, overlap_length, gaps, mismatches)."""
    if args.paired:
        assert reverse_sequence and reverse_quals
        reverse_sequence, reverse_quals = reverse_complement(reverse_sequence, reverse_quals)
        forward_aligned, reverse_aligned, _score = nw_align(forward_sequence, reverse_sequence, args.match, args.mismatch, args.gap)
        contig, contig_quals, overlap_length, gaps, mismatches = make_contig(
            forward_aligned, forward_quals, reverse_aligned, reverse_quals, args.insert, args.deltaq, args.consensus_qscore,
            args.qscore_cap, args.trim_overlap)
    else:
        overlap_length, gaps, mismatches = 0, 0, 0
        contig, contig_quals = forward_sequence, forward_quals
    if args.truncate:
        contig, contig_quals = contig[:args.truncate], contig_quals[:args.truncate]
    if args.only_contig:
        expected_errors = 0
    else:
        contig_quals = [qual if qual > 0 else 1 for qual in contig_quals]               # moira.py:814
        if args.error_calc in ("poisson_binomial", "poisson_binomial_py"):
            if args.error_calc == "poisson_binomial":
                from . import bernoulli                                                   # the compiled module's rule: 'N' and 'n'
                expected_errors, Ns = bernoulli.calculate_errors_PB(contig, contig_quals, args.alpha)
            else:
                expected_errors, Ns = calculate_errors_PB(contig, contig_quals, args.alpha)
        elif args.error_calc == "poisson":
            expected_errors, Ns = calculate_errors_poisson(contig, contig_quals, args.alpha)
        else:
            raise ValueError("--error_calc bootstrap is not offered (deprecated upstream, non-deterministic)")
        if args.ambigs == "treat_as_errors":
            expected_errors = expected_errors + Ns
    if args.round:
        expected_errors = math.floor(expected_errors)
    return header, contig, contig_quals, expected_errors, overlap_length, gaps, mismatches


__all__ = ["calculate_errors_PB", "calculate_errors_poisson", "interpolate", "reverse_complement", "nw_align", "make_contig",
           "process_data", "parse_arguments", "check_arguments", "main", "open_input", "LengthMismatchError", "NameMismatchError",
           "EmptySeqError", "EmptyQualError", "UnpairedFilesError", "ReturnedNaNError"]
