"""CPU: pins the oracle (oracle/pb_oracle.c, oracle/py_oracle.py) against the reference's own
known-answer tests, golden partitions and outputs of the unmodified reference binary."""
import numpy as np
import pytest

from oracle import py_oracle as po


def _q1(kat):
    return [ord(c) - kat["fastq_offset"] for c in kat["testQual1_ascii"]]


def test_pb_kat_both_variants(kat):
    # moira/test/test_moira.py:39-43 -- exact float equality
    exp = tuple(kat["pb_expected"])
    assert po.pb_c(kat["testSeq1"], _q1(kat), kat["alpha"]) == exp
    assert po.pb_c(kat["testSeq1"], _q1(kat), kat["alpha"], faithful=True) == exp


def test_poisson_kat(kat):
    # moira/test/test_moira.py:44-45
    assert po.calculate_errors_poisson(kat["testSeq1"], _q1(kat), kat["alpha"]) == tuple(kat["poisson_expected"])


def test_process_forward_kat(kat):
    # moira/test/test_moira.py:63-66, 127: truncate=200 then PB (+Ns)
    args = po.Args(truncate=200)
    contig, quals, ee = po.process_filter(kat["testSeq1"], _q1(kat), args)
    fp = kat["forward_process"]
    assert contig == fp["seq"] and quals == fp["quals"] and ee == fp["ee"]


def test_process_paired_kat_contig(kat):
    # moira/test/test_moira.py:67-70, 128: the listed contig + qualities give the listed ee
    pp = kat["paired_process"]
    _, _, ee = po.process_filter(pp["seq"], pp["quals"], po.Args())
    assert ee == pp["ee"]


def test_forward_pipeline_partition(forward_records, forward_names):
    # moira/test/test_moira.py:74-87: 1000 reads -> 122 good / 365 bad uniques, same member lists
    good, bad = po.collapse_and_decide(forward_records, po.Args())
    assert len(good) == 122 and len(bad) == 365
    assert {k: v[0] for k, v in good.items()} == forward_names["good"]
    assert {k: v[0] for k, v in bad.items()} == forward_names["bad"]
    assert all(v[2] == po.REASON_ERRORS for v in bad.values())


def test_golden_contigs_reclassify(contigs):
    # moira/test/test_results/paired.qc.*: 324 good / 76 bad real 253-bp contigs
    args = po.Args()
    n_good = 0
    for c in contigs:
        _, _, ee = po.process_filter(c["seq"], c["quals"], args)
        ok, _ = po.decide(c["seq"], ee, args)
        assert ok == (c["label"] == "good"), c["header"]
        n_good += ok
    assert n_good == 324 and len(contigs) == 400


def test_oracle_matches_reference_binary_outputs(forward_records, contigs, ref_outputs):
    # bit-exact against outputs of oracle/_ref (the unmodified reference), committed in tests/golden
    ee = np.array([po.pb_c(s, [q if q > 0 else 1 for q in ql], 0.005)[0] for _, s, ql in forward_records])
    assert np.array_equal(ee, ref_outputs["forward_ee"])
    cee = np.array([po.pb_c(c["seq"], c["quals"], 0.005)[0] for c in contigs])
    assert np.array_equal(cee, ref_outputs["contigs_ee"])


def test_oracle_batch_matches_reference_on_synthetic(ref_outputs):
    slab, off, ln = ref_outputs["syn_slab"], ref_outputs["syn_offsets"], ref_outputs["syn_lengths"]
    alphas = ref_outputs["syn_alpha"]
    for a in np.unique(alphas):
        sel = np.nonzero(alphas == a)[0]
        ee, ns = po.pb_batch(slab, off[sel], ln[sel], float(a))
        assert np.array_equal(ee, ref_outputs["syn_ee"][sel])
        assert np.array_equal(ns, ref_outputs["syn_ns"][sel])
    # the faithful O(L j*^2) variant on a subset
    sel = np.arange(0, 4000, 16)
    for a in np.unique(alphas[sel]):
        s2 = sel[alphas[sel] == a]
        ee, ns = po.pb_batch(slab, off[s2], ln[s2], float(a), faithful=True)
        assert np.array_equal(ee, ref_outputs["syn_ee"][s2])


@pytest.mark.skipif(not po.have_ref(), reason="oracle/_ref not built (reference tree absent)")
def test_oracle_matches_live_reference(kat):
    ref = po.ref_module()
    assert ref.calculate_errors_PB(kat["testSeq1"], _q1(kat), 0.005) == tuple(kat["pb_expected"])
    from moira_b200 import synth
    slab, off, ln = synth.generate("v4", 300, 7)
    ee_r, ns_r = po.ref_batch(slab, off, ln, 0.005, n_threads=2)
    ee_o, ns_o = po.pb_batch(slab, off, ln, 0.005)
    assert np.array_equal(ee_r, ee_o) and np.array_equal(ns_r, ns_o)
    # 1500-bp reads need the big-stack worker threads
    slab, off, ln = synth.generate("ccs", 6, 11)
    ee_r, _ = po.ref_batch(slab, off, ln, 0.005, n_threads=2)
    ee_o, _ = po.pb_batch(slab, off, ln, 0.005)
    assert np.array_equal(ee_r, ee_o)


def test_edge_cases():
    assert po.pb_c("", [], 0.005) == (0.0, 0)                       # L' == 0 -> 0 (bernoullimodule.c:257-260)
    assert po.pb_c("NNnN", [2, 2, 2, 2], 0.005) == (0.0, 4)         # all-N (patch-notes.md:50-51)
    assert po.pb_c("A", [40], 0.005) == (0.0, 0)                    # j* == 0 -> interpolation clamps to 0
    ee, ns = po.pb_c("A", [3], 0.005)                               # single bad base: j* == 1
    assert 0.0 < ee <= 1.0 and ns == 0
    assert po.pb_c("AC", [0, 0], 0.3) == po.pb_c("AC", [1, 1], 0.3)  # Q == 0 -> 1 (bernoullimodule.c:104-107)


def test_decide_batch_matches_scalar():
    rng = np.random.default_rng(3)
    from moira_b200 import synth
    slab, off, ln = synth.generate("mixed", 200, 5)
    ee, ns = po.pb_batch(slab, off, ln, 0.005)
    has_n = np.array([(slab[int(o):int(o) + int(l)] == 0xFF).any() for o, l in zip(off, ln)])
    for kw in (dict(thr_kind="uncert", thr=0.01, ambigs="treat_as_errors", round_flag=False, truncate=None),
               dict(thr_kind="maxerrors", thr=2.0, ambigs="disallow", round_flag=True, truncate=250)):
        ok, reason, eef = po.decide_batch(ee, ns, ln, has_n, **kw)
        args = po.Args(uncert=kw["thr"] if kw["thr_kind"] == "uncert" else None,
                       maxerrors=kw["thr"] if kw["thr_kind"] == "maxerrors" else None,
                       ambigs=kw["ambigs"], round=kw["round_flag"], truncate=kw["truncate"])
        for i in rng.choice(200, 40, replace=False):
            seq = "N" * int(has_n[i]) + "A" * (int(ln[i]) - int(has_n[i]))
            assert (bool(ok[i]), int(reason[i])) == po.decide(seq, float(eef[i]), args)


def test_property_restatements_agree_with_reference_binary():
    """hypothesis: random short reads (N/n, Q = 0..93, any alpha): the O(L j*) restatement, the faithful
    O(L j*^2) restatement and -- when built -- the unmodified reference binary agree bit for bit."""
    from hypothesis import given, settings, strategies as st

    ref = po.ref_module() if po.have_ref() else None
    read = st.lists(st.tuples(st.sampled_from("ACGTNn"), st.integers(0, 93)), min_size=0, max_size=60)

    @settings(max_examples=300, deadline=None)
    @given(read, st.sampled_from([0.5, 0.2, 0.05, 0.005, 0.001, 1e-6]))
    def check(pairs, alpha):
        seq = "".join(b for b, _ in pairs)
        quals = [q for _, q in pairs]
        fast = po.pb_c(seq, quals, alpha)
        assert fast == po.pb_c(seq, quals, alpha, faithful=True)
        if ref is not None:
            assert fast == ref.calculate_errors_PB(seq, quals, alpha)
        assert fast[1] == seq.count("N") + seq.count("n") and fast[0] >= 0.0

    check()


# ---- paired-end contig constructor (oracle/contig_oracle.c) ------------------------------------------
def _q2(kat):
    return [ord(c) - kat["fastq_offset"] for c in kat["testQual2_ascii"]]


def test_contig_constructor_kats(kat):
    # moira/test/test_moira.py:49-59 -- testReverseComplement, testNwPython / testNwC, testContig
    a = kat["contig_args"]
    rc_seq, rc_q = po.reverse_complement(kat["testSeq2"], _q2(kat))
    assert [rc_seq, rc_q] == kat["testRC2"]
    al = po.nw_align(kat["testSeq1"], rc_seq, a["match"], a["mismatch"], a["gap"])
    assert list(al) == kat["test_aligned"]
    ct = po.make_contig(al[0], _q1(kat), al[1], rc_q, a["insert"], a["deltaq"], a["consensus_qscore"], a["qscore_cap"],
                        a["trim_overlap"])
    assert list(ct) == kat["test_contig"]
    # process_data, paired (test_moira.py:67-70): contig -> truncate 200 -> PB
    contig, quals, ov, gaps, mism = po.pair_to_contig(kat["testSeq1"], _q1(kat), kat["testSeq2"], _q2(kat))
    pp = kat["paired_process"]
    c200, q200, ee = po.process_filter(contig, quals, po.Args(truncate=200))
    assert (c200, q200, ee, ov, gaps, mism) == (pp["seq"], pp["quals"], pp["ee"], pp["overlap"], pp["gaps"], pp["mismatches"])
    with pytest.raises(ValueError):
        po.reverse_complement("ACGX")


def test_aligner_matches_reference_aligner_outputs(ref_alignments):
    # committed outputs of the unmodified reference aligner (oracle/_ref/nw_align.so)
    for c in ref_alignments:
        got = po.nw_align(c["seq_1"], c["seq_2"], c["match"], c["mismatch"], c["gap"])
        assert got == (c["aligned_1"], c["aligned_2"], c["score"])


@pytest.mark.skipif(not po.have_ref_nw(), reason="oracle/_ref/nw_align.so not built (reference tree absent)")
def test_aligner_matches_live_reference_aligner(kat):
    nw = po.ref_nw_module()
    rng = np.random.default_rng(99)
    for it in range(150):
        a = "".join(rng.choice(list("ACGTN"), int(rng.integers(1, 90))))
        b = "".join(rng.choice(list("ACGTN"), int(rng.integers(1, 90))))
        if it % 2:
            b = a[int(rng.integers(0, len(a))):] + b
        m, x, g = [(1, -1, -2), (3, -2, -4), (1, 0, -1)][it % 3]
        assert po.nw_align(a, b, m, x, g) == tuple(nw.nw_align(a, b, m, x, g))
    rc = po.reverse_complement(kat["testSeq2"])
    assert list(nw.nw_align(kat["testSeq1"], rc, 1, -1, -2)) == kat["test_aligned"]


def test_paired_pipeline_contigs_and_partition(oracle_contigs, contigs, paired_names):
    # moira/test/test_moira.py:88-101: test1 + test2 -> 1000 contigs -> collapse -> golden fasta / qual / names
    by_header = {h: (c, q) for h, c, q, _, _, _ in oracle_contigs}
    for g in contigs:                                   # every golden representative is its pair's contig
        c, q = by_header[g["header"]]
        assert c == g["seq"] and [v if v > 0 else 1 for v in q] == g["quals"]
    good, bad = po.collapse_and_decide([(h, c, q) for h, c, q, _, _, _ in oracle_contigs], po.Args())
    assert {k: v[0] for k, v in good.items()} == paired_names["good"]
    assert {k: v[0] for k, v in bad.items()} == paired_names["bad"]


def test_make_contig_modes_and_validation():
    fa, ra = "ACGTAC--", "--GTTCGA"
    fq, rq = [30, 30, 30, 10, 12, 12], [20, 20, 14, 40, 25, 25]
    best = po.make_contig(fa, fq, ra, rq, 20, 6, "best", 40, False)
    assert best == ("ACGTNCGA", [30, 30, 30, 20, 2, 40, 25, 25], 3, 0, 1)          # |12-14| < 6 -> N, Q 2
    assert po.make_contig(fa, fq, ra, rq, 20, 6, "sum", 0, False)[1] == [30, 30, 50, 30, 2, 52, 25, 25]
    assert po.make_contig(fa, fq, ra, rq, 20, 6, "sum", 40, True) == ("GTNC", [40, 30, 2, 40], 3, 0, 1)
    post = po.make_contig(fa, fq, ra, rq, 20, 6, "posterior", 0, False)
    assert post[0] == "ACGTTCGA" and post[1][2] == 54 and post[1][4] == 4           # Edgar & Flyvbjerg posteriors
    gap = po.make_contig("ACG-TT", [30, 30, 30, 30, 30], "ACGATT", [30, 30, 30, 21, 30, 30], 20, 6, "best", 40, False)
    assert gap == ("ACGATT", [30, 30, 30, 21, 30, 30], 5, 1, 0)                     # inserted: 21 > insert
    gap = po.make_contig("ACG-TT", [30, 30, 30, 30, 30], "ACGATT", [30, 30, 30, 20, 30, 30], 20, 6, "best", 40, False)
    assert gap[0] == "ACGTT" and gap[3] == 1                                         # 20 is not > insert: dropped
    assert po.make_contig("ACG-TT", [30] * 5, "ACGATT", [30] * 6, 20, 6, "posterior", 40, False)[0] == "ACGNTT"
    with pytest.raises(ValueError):
        po.make_contig("ACG", [30, 30], "ACG", [30, 30, 30], 20, 6, "best", 40, False)
    with pytest.raises(ValueError):
        po.make_contig("ACG", [30] * 3, "ACG", [30] * 3, 0, 6, "best", 40, False)
