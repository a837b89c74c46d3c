"""Drop-in for moira's compiled `nw_align` module (moira/nw_align.pyx): `import nw_align` with this repository's
root on sys.path gives nw_align.nw_align(seq_1, seq_2, match, mismatch, gap) computed by the CUDA path."""
from moira_b200.contig import nw_align  # noqa: F401
