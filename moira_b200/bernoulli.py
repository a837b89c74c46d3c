"""Drop-in for the reference's compiled `bernoulli` extension module
(/root/reference/moira/bernoullimodule.c:66-125): same function name, argument meaning, return
value and exception types, computed by the sm_100a CUDA path as a batch of one.

    import moira_b200.bernoulli as bernoulli          # or: `import bernoulli` with the repo root on sys.path
    ee, Ns = bernoulli.calculate_errors_PB(contig, contig_quals, alpha)     # moira.py:817, test_moira.py:43
"""
from __future__ import annotations

from . import _lib as L
from .api import Context

__doc__ = ("This module provides an interface for calculating the expected errors of a given sequence "
           "using a sum of Bernoulli random variables.")

_ctx = None


def _context() -> Context:
    global _ctx
    if _ctx is None:
        _ctx = Context(0)
    return _ctx


def calculate_errors_PB(contig, contig_quals, alpha):
    """This function returns the expected errors of a given sequence with a given confidence value
    using a sum of Bernoulli random variables.

    Mirrors PyArg_ParseTuple("sO!d") + the checks of bernoullimodule.c:74-108: `contig` must be a
    str, `contig_quals` a list of ints (TypeError otherwise), 0 < alpha < 1 and equal lengths
    (ValueError with the reference's messages)."""
    if not isinstance(contig, str):
        raise TypeError("argument 1 must be str, not %s" % type(contig).__name__)
    if not isinstance(contig_quals, list):
        raise TypeError("argument 2 must be list, not %s" % type(contig_quals).__name__)
    try:
        alpha = float(alpha)
    except (TypeError, ValueError):
        raise TypeError("must be real number, not %s" % type(alpha).__name__)
    if alpha <= 0 or alpha >= 1:
        raise ValueError("Alpha must be between 0 and 1")                              # bernoullimodule.c:81
    if len(contig_quals) != len(contig):
        raise ValueError("contig and contig_quals must have the same length")          # bernoullimodule.c:88
    for q in contig_quals:
        if not isinstance(q, int):
            raise TypeError("an integer is required (got type %s)" % type(q).__name__)  # PyLong_AsLong, :97-102
    try:
        return _context().calculate_errors_PB(contig, contig_quals, alpha)
    except L.MoiraError as exc:
        if exc.code in (L.ERR_BAD_ALPHA, L.ERR_LENGTH_MISMATCH, L.ERR_BAD_QUALITY):
            raise ValueError(exc.message) from None
        raise
