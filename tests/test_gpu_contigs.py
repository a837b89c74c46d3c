"""GPU: the paired-end contig constructor (moira_contig.cu through the C ABI) against the oracle
(oracle/contig_oracle.c), the reference's known-answer vectors and the committed outputs of the unmodified
reference aligner."""
import numpy as np
import pytest

import moira_b200
from moira_b200 import ContigParams, FilterParams
from moira_b200 import _lib as L
from oracle import py_oracle as po

pytestmark = pytest.mark.gpu


def _q(kat, key):
    return [ord(c) - kat["fastq_offset"] for c in kat[key]]


def test_contig_kats_through_the_reference_named_functions(kat):
    # moira/test/test_moira.py:49-59 with this repo's `nw_align` / contig modules in place of the reference's
    import nw_align
    from moira_b200 import contig
    a = kat["contig_args"]
    rc_seq, rc_q = contig.reverse_complement(kat["testSeq2"], _q(kat, "testQual2_ascii"))
    assert [rc_seq, rc_q] == kat["testRC2"]
    al = nw_align.nw_align(kat["testSeq1"], rc_seq, a["match"], a["mismatch"], a["gap"])
    assert list(al) == kat["test_aligned"]
    ct = contig.make_contig(al[0], _q(kat, "testQual1_ascii"), al[1], rc_q, a["insert"], a["deltaq"], a["consensus_qscore"],
                            a["qscore_cap"], a["trim_overlap"])
    assert list(ct) == kat["test_contig"]
    with pytest.raises(ValueError):
        contig.reverse_complement("ACGU")
    with pytest.raises(contig.LengthMismatchError):
        contig.make_contig("ACG", [30, 30], "ACG", [30, 30, 30], 20, 6, "best", 40, False)
    with pytest.raises(ValueError):
        contig.make_contig("ACG", [30] * 3, "ACG", [30] * 3, 0, 6, "best", 40, False)
    with pytest.raises(ValueError):
        contig.make_contig("ACG", [30] * 3, "ACG", [30] * 3, 20, 6, "worst", 40, False)


def test_single_alignments_match_the_reference_aligner_outputs(ctx, ref_alignments):
    for c in ref_alignments:
        got = ctx.nw_align(c["seq_1"], c["seq_2"], c["match"], c["mismatch"], c["gap"])
        assert got == (c["aligned_1"], c["aligned_2"], c["score"]), (c["seq_1"], c["seq_2"])


def test_make_contig_modes_match_oracle(ctx):
    fa, ra = "ACGTAC--", "--GTTCGA"
    fq, rq = [30, 30, 30, 10, 12, 12], [20, 20, 14, 40, 25, 25]
    for mode in ("best", "sum", "posterior"):
        for cap in (0, 40):
            for trim in (False, True):
                p = ContigParams(consensus_qscore=mode, qscore_cap=cap, trim_overlap=trim)
                assert ctx.make_contig(fa, fq, ra, rq, p) == po.make_contig(fa, fq, ra, rq, 20, 6, mode, cap, trim)
    for rq3 in (20, 21):
        a = ("ACG-TT", [30] * 5, "ACGATT", [30, 30, 30, rq3, 30, 30])
        assert ctx.make_contig(*a, ContigParams()) == po.make_contig(*a, 20, 6, "best", 40, False)


def _fixture_pairs(forward_records, reverse_records):
    fs, fq, fo, fl = moira_b200.pack_sequences([r[1] for r in forward_records], [r[2] for r in forward_records])
    rs, rq, ro, rl = moira_b200.pack_sequences([r[1] for r in reverse_records], [r[2] for r in reverse_records])
    return fs, fq, fo, fl, rs, rq, ro, rl


def test_fixture_pairs_to_contigs_and_filter(ctx, forward_records, reverse_records, oracle_contigs):
    """The reference's 1000 MiSeq pairs (test1/test2.fastq): contigs, overlap statistics and the filter's
    expected errors on the contigs, all from one moira_filter_pairs call, equal the oracle's."""
    args = _fixture_pairs(forward_records, reverse_records)
    res = ctx.filter_pairs(*args, ContigParams(), FilterParams(exact_ee=True))
    assert not res.status.any()
    for r, (_, contig, quals, ov, gaps, mism) in enumerate(oracle_contigs):
        assert res.contig(r) == (contig, quals), r
        assert (int(res.overlap[r]), int(res.gaps[r]), int(res.mismatches[r])) == (ov, gaps, mism)
    ee = np.array([po.pb_c(c, [q if q > 0 else 1 for q in ql], 0.005)[0] for _, c, ql, _, _, _ in oracle_contigs])
    ns = np.array([c.count("N") + c.count("n") for _, c, _, _, _, _ in oracle_contigs])
    assert np.array_equal(res.filter.ee, ee) and np.array_equal(res.filter.ns, ns)
    ok, _, _ = po.decide_batch(ee, ns, res.contig_len, ns > 0, thr_kind="uncert", thr=0.01, ambigs="treat_as_errors",
                               round_flag=False, truncate=None)
    assert np.array_equal((res.filter.flags & L.FLAG_ACCEPT) != 0, ok)
    assert int(res.filter.counters[L.CNT_READS]) == 1000 and int(res.filter.counters[L.CNT_ACCEPTED]) == int(ok.sum())
    # contigs only (no filter), and straight from the FASTQ text: no repacking on the host
    only = ctx.filter_pairs(*args, ContigParams())
    assert only.filter is None and np.array_equal(only.contig_len, res.contig_len)
    assert all(only.contig(r) == res.contig(r) for r in range(0, 1000, 37))


def _random_pairs(rng, n, lo, hi, n_rate=0.01):
    fwd, rev = [], []
    comp = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N", "R": "Y", "Y": "R"}
    for k in range(n):
        l1, l2 = int(rng.integers(lo, hi + 1)), int(rng.integers(lo, hi + 1))
        frag = "".join(rng.choice(list("ACGT"), l1 + l2))
        ov = int(rng.integers(0, min(l1, l2) + 1)) if k % 5 else 0
        f = list(frag[:l1])
        start = max(0, l1 - ov)
        r_fwd_strand = list(frag[start:start + l2])
        r_fwd_strand += list(rng.choice(list("ACGT"), l2 - len(r_fwd_strand)))
        for s in (f, r_fwd_strand):                     # substitutions, ambiguity codes, one indel now and then
            for i in range(len(s)):
                u = rng.random()
                if u < 0.03:
                    s[i] = "ACGT"[int(rng.integers(4))]
                elif u < 0.03 + n_rate:
                    s[i] = "NRY"[int(rng.integers(3))]
            if rng.random() < 0.3 and len(s) > 3:
                i = int(rng.integers(len(s)))
                if rng.random() < 0.5:
                    del s[i]
                else:
                    s.insert(i, "ACGT"[int(rng.integers(4))])
        f = f[:hi] or ["A"]
        r_fwd_strand = r_fwd_strand[:hi] or ["C"]
        r = [comp[c] for c in reversed(r_fwd_strand)]
        fq = [int(v) for v in rng.choice([2, 8, 12, 15, 20, 21, 25, 30, 33, 37, 38, 40, 0], len(f))]
        rq = [int(v) for v in rng.choice([2, 8, 12, 15, 20, 21, 25, 30, 33, 37, 38, 40, 93], len(r))]
        fwd.append(("p%d" % k, "".join(f), fq))
        rev.append(("p%d" % k, "".join(r), rq))
    return fwd, rev


@pytest.mark.parametrize("lo,hi,n", [(1, 40, 600), (60, 128, 300), (129, 256, 200), (257, 320, 120), (321, 512, 60), (513, 1024, 24)])
def test_random_pairs_every_strip_width_and_option(ctx, lo, hi, n):
    """Each length class runs another columns-per-lane instantiation (4, 8, 10, 16, 32); options cycle through
    consensus modes, caps, trimming and scoring schemes."""
    rng = np.random.default_rng(hi)
    fwd, rev = _random_pairs(rng, n, lo, hi)
    args = _fixture_pairs(fwd, rev)
    variants = [ContigParams(), ContigParams(consensus_qscore="sum", qscore_cap=0), ContigParams(consensus_qscore="posterior"),
                ContigParams(trim_overlap=True, insert=10, deltaq=3), ContigParams(match=2, mismatch=-3, gap=-5, qscore_cap=35),
                ContigParams(match=1, mismatch=-1, gap=-1, consensus_qscore="posterior", qscore_cap=0, trim_overlap=True)]
    for vi, p in enumerate(variants if n >= 100 else variants[:3]):
        res = ctx.filter_pairs(*args, p, FilterParams(exact_ee=True) if vi == 0 else None)
        assert not res.status.any()
        for r in range(n):
            want = po.pair_to_contig(fwd[r][1], fwd[r][2], rev[r][1], rev[r][2], p.match, p.mismatch, p.gap, p.insert, p.deltaq,
                                     p.consensus_qscore, p.qscore_cap, p.trim_overlap)
            assert res.contig(r) == (want[0], want[1]), (vi, r)
            assert (int(res.overlap[r]), int(res.gaps[r]), int(res.mismatches[r])) == want[2:], (vi, r)
        if vi == 0:
            for r in range(0, n, 7):
                c, ql = res.contig(r)
                assert res.filter.ee[r] == po.pb_c(c, [q if q > 0 else 1 for q in ql], 0.005)[0]


def test_pair_status_codes_and_argument_errors(ctx):
    fwd = [("a", "ACGTACGT", [30] * 8), ("b", "ACGT", [30] * 4), ("c", "ACGTAC", [30] * 6), ("d", "ACGT", [30] * 4)]
    rev = [("a", "ACGTACGT", [30] * 8), ("b", "ACXT", [30] * 4), ("c", "GTACGT", [250] * 6), ("d", "ACGT", [30] * 4)]
    fs, fq, fo, fl = moira_b200.pack_sequences([r[1] for r in fwd], [r[2] for r in fwd])
    rs, rq, ro, rl = moira_b200.pack_sequences([r[1] for r in rev], [r[2] for r in rev])
    rl = rl.copy()
    rl[3] = 0
    res = ctx.filter_pairs(fs, fq, fo, fl, rs, rq, ro, rl, ContigParams(consensus_qscore="sum", qscore_cap=0), FilterParams())
    assert list(res.status) == [L.PAIR_OK, L.PAIR_BAD_BASE, L.PAIR_BAD_QUALITY, L.PAIR_EMPTY]
    assert res.contig_len[1] == 0 and res.contig_len[3] == 0 and res.contig(0)[0] == "ACGTACGT"
    assert res.filter.ee[1] == 0.0 and int(res.filter.counters[L.CNT_READS]) == 4
    long_rev = moira_b200.pack_sequences(["ACGT" * 300], [[30] * 1200])
    one_fwd = moira_b200.pack_sequences(["ACGT"], [[30] * 4])
    res = ctx.filter_pairs(*one_fwd, *long_rev, ContigParams())
    assert list(res.status) == [L.PAIR_TOO_LONG]
    with pytest.raises(moira_b200.MoiraError) as ei:
        ctx.filter_pairs(fs, fq, fo, fl, rs, rq, ro, rl, ContigParams(insert=0))
    assert ei.value.code == L.ERR_BAD_ARG and "insert" in ei.value.message
    assert ctx.nw_align("", "ACG", 1, -1, -2) == ("---", "ACG", 0)


def test_pairs_straight_from_fastq_text(ctx, oracle_contigs):
    """The zero-repack path: both FASTQ texts are handed over as they are, with the parser's offsets."""
    import bz2, gzip, os
    from conftest import GOLDEN
    t1 = gzip.open(os.path.join(GOLDEN, "test1.fastq.gz"), "rb").read()
    t2 = bz2.open(os.path.join(GOLDEN, "test2.fastq.bz2"), "rb").read()
    _, _, l1, _, _, s1, q1 = moira_b200.parse_fastq(t1, 33, True)
    _, _, l2, _, _, s2, q2 = moira_b200.parse_fastq(t2, 33, True)
    b1, b2 = np.frombuffer(t1, np.uint8), np.frombuffer(t2, np.uint8)
    res = ctx.filter_pairs(b1, b1, s1, l1, b2, b2, s2, l2, ContigParams(), None, True, q1, q2, 33)
    assert not res.status.any()
    for r in (0, 1, 500, 999):
        assert res.contig(r) == (oracle_contigs[r][1], oracle_contigs[r][2])
    big = 40                                                # the batch repeated: several chunks on both streams
    n = len(l1)
    res2 = ctx.filter_pairs(b1, b1, np.tile(s1, big), np.tile(l1, big), b2, b2, np.tile(s2, big), np.tile(l2, big), ContigParams(),
                            FilterParams(exact_ee=False), True, np.tile(q1, big), np.tile(q2, big), 33)
    assert np.array_equal(res2.contig_len, np.tile(res.contig_len, big))
    valid = np.arange(res.contig_seq.shape[1])[None, :] < res.contig_len[:, None]     # bytes past a contig's end are not defined
    for rep in (0, big // 2, big - 1):                      # chunks handled by either stream
        assert np.array_equal(np.where(valid, res2.contig_seq.reshape(big, n, -1)[rep], 0), np.where(valid, res.contig_seq, 0))
        assert np.array_equal(np.where(valid, res2.contig_qual.reshape(big, n, -1)[rep], 0), np.where(valid, res.contig_qual, 0))
    assert int(res2.filter.counters[L.CNT_READS]) == big * n
