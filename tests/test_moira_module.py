"""The reference's own test file (moira/test/test_moira.py), test by test, against `import moira` of this repository
(top-level moira.py -> moira_b200/reference_api.py): same function names, argument order, expected values.  Vectors:
tests/golden/kat.json (copied from that file by tests/golden/make_golden.py).  GPU: every call computes on the device."""
import bz2
import gzip
import os

import pytest

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class Arguments:
    """test_moira.py:33-36."""

    def __init__(self, **kwargs):
        for key, value in kwargs.items():
            setattr(self, key, value)


def _args(**over):
    """test_moira.py:130-137."""
    kw = dict(alpha=0.005, match=1, gap=-2, mismatch=-1, insert=20, deltaq=6, consensus_qscore="best", paired=True, truncate=200,
              only_contig=False, error_calc="poisson_binomial", ambigs="treat_as_errors", round=False, silent=True, nowarnings=False,
              doc=False, uncert=0.01, maxerrors=None, processors=1, forward_fasta=None, forward_quals=None, reverse_fasta=None,
              reverse_quals=None, forward_fastq=None, reverse_fastq=None, output_format="fasta", collapse=True, pipeline="mothur",
              fastq_offset=33, relabel=None, output_compression="none", qscore_cap=40, min_overlap=None, trim_overlap=False)
    kw.update(over)
    return Arguments(**kw)


@pytest.fixture(scope="module")
def v(kat):
    q = lambda key: [ord(c) - 33 for c in kat[key]]   # noqa: E731  (test_moira.py:124-125)
    return dict(testSeq1=kat["testSeq1"], testSeq2=kat["testSeq2"], testQual1=q("testQual1_ascii"), testQual2=q("testQual2_ascii"),
                testRC2=tuple(kat["testRC2"]), test_aligned=tuple(kat["test_aligned"]), test_contig=tuple(kat["test_contig"]), kat=kat)


# ---- class TestErrorCalculationAlgorithms (test_moira.py:38-45) ----------------------------------------------------------
def test_PbPython(v):
    import moira
    assert moira.calculate_errors_PB(v["testSeq1"], v["testQual1"], 0.005) == (6.446879136706666, 0)


def test_PbC(v):
    import bernoulli
    assert bernoulli.calculate_errors_PB(v["testSeq1"], v["testQual1"], 0.005) == (6.446879136706666, 0)


def test_Poisson(v):
    import moira
    ee, ns = moira.calculate_errors_poisson(v["testSeq1"], v["testQual1"], 0.005)
    assert ns == 0 and abs(ee - 6.932519986616133) <= 1e-12 * 6.932519986616133   # device exp / pow: 1e-12 relative


# ---- class TestContigConstructor (test_moira.py:48-60) --------------------------------------------------------------------
def test_ReverseComplement(v):
    import moira
    assert moira.reverse_complement(v["testSeq2"], v["testQual2"]) == v["testRC2"]


def test_NwPython_and_NwC(v):
    import moira
    import nw_align
    a = _args()
    want = v["test_aligned"]
    assert moira.nw_align(v["testSeq1"], moira.reverse_complement(v["testSeq2"]), a.match, a.mismatch, a.gap) == want
    assert nw_align.nw_align(v["testSeq1"], moira.reverse_complement(v["testSeq2"]), a.match, a.mismatch, a.gap) == want


def test_Contig(v):
    import moira
    a = _args()
    testSeq2RC, testQual2RC = moira.reverse_complement(v["testSeq2"], v["testQual2"])
    aligned1, aligned2 = moira.nw_align(v["testSeq1"], testSeq2RC, a.match, a.mismatch, a.gap)[:2]
    got = moira.make_contig(aligned1, v["testQual1"], aligned2, testQual2RC, a.insert, a.deltaq, a.consensus_qscore, a.qscore_cap,
                            a.trim_overlap)
    assert tuple(got) == v["test_contig"]


# ---- class TestProcessing (test_moira.py:63-70) ------------------------------------------------------------------------------
def test_ProcessForwardSeq(v):
    import moira
    fp = v["kat"]["forward_process"]
    got = moira.process_data("foo", v["testSeq1"], v["testQual1"], None, None, _args(truncate=200, paired=False))
    assert got == ("foo", fp["seq"], fp["quals"], fp["ee"], 0, 0, 0)


def test_ProcessPairedSeq(v):
    import moira
    pp = v["kat"]["paired_process"]
    got = moira.process_data("foo", v["testSeq1"], v["testQual1"], v["testSeq2"], v["testQual2"], _args(truncate=200, paired=True))
    assert got == ("foo", pp["seq"], pp["quals"], pp["ee"], pp["overlap"], pp["gaps"], pp["mismatches"])


# ---- class TestFullPipeline (test_moira.py:73-113): moira.main(args) with a hand-made argument object ---------------------
def _read(path):
    opener = gzip.open if path.endswith(".gz") else bz2.open if path.endswith(".bz2") else open
    with opener(path, "rb") as fh:
        return fh.read()


def test_ProcessForwardDataset_PairedDataset_Compression(tmp_path):
    """moira.main(Arguments(...)) writes the files the command line writes (whose records tests/test_cli.py compares with the
    reference's test_results/), for the forward run, the paired run, and compressed inputs / outputs."""
    import moira
    from moira_b200 import cli
    f1, f2 = os.path.join(GOLDEN, "test1.fastq.gz"), os.path.join(GOLDEN, "test2.fastq.bz2")
    suffixes = [".qc.good.fasta", ".qc.good.qual", ".qc.good.names", ".qc.bad.fasta", ".qc.bad.qual", ".qc.bad.names"]
    # forward
    a = _args(truncate=None, paired=False, forward_fastq=f1, reverse_fastq=None, output_prefix=str(tmp_path / "forward"))
    assert moira.main(a) == 0
    assert cli.run(["-ffq", f1, "-op", str(tmp_path / "forward_cli"), "--silent"]) == 0
    for s in suffixes:
        assert _read(str(tmp_path / "forward") + s) == _read(str(tmp_path / "forward_cli") + s), s
    # paired, then compressed outputs
    a = _args(truncate=None, paired=True, forward_fastq=f1, reverse_fastq=f2, output_prefix=str(tmp_path / "paired"))
    assert moira.main(a) == 0
    assert cli.run(["-ffq", f1, "-rfq", f2, "--paired", "-op", str(tmp_path / "paired_cli"), "--silent"]) == 0
    for s in suffixes:
        assert _read(str(tmp_path / "paired") + s) == _read(str(tmp_path / "paired_cli") + s), s
    for comp in ("gz", "bz2"):
        a = _args(truncate=None, paired=True, forward_fastq=f1, reverse_fastq=f2, output_prefix=str(tmp_path / "pc"), output_compression=comp)
        assert moira.main(a) == 0
        assert _read("%s.qc.good.fasta.%s" % (str(tmp_path / "pc"), comp)) == _read(str(tmp_path / "paired") + ".qc.good.fasta")
