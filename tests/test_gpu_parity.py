"""GPU (-m gpu): the CUDA path, called through the C ABI, against the oracle and the committed
golden vectors.  Bit-exact for Poisson-binomial ee / Ns / decisions; the Poisson mode within the
stated FP64 tolerance (device exp/pow vs glibc), decisions identical outside the 1e-12 band."""
import gzip
import os

import numpy as np
import pytest

import moira_b200
from moira_b200 import FilterParams, synth
from moira_b200 import _lib as L
from oracle import py_oracle as po

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _q1(kat):
    return [ord(c) - kat["fastq_offset"] for c in kat["testQual1_ascii"]]


def _same_counters(a, b):
    """Every counter but MOIRA_CNT_ESCALATED, which describes the route a batch took (cascade or single sweep; the
    library adapts that to the data it has seen), not its result."""
    keep = (np.arange(L.N_COUNTERS) != L.CNT_ESCALATED) & (np.arange(L.N_COUNTERS) != L.CNT_FP64_OPS) & (np.arange(L.N_COUNTERS) != L.CNT_CLASSIFIED)
    return np.array_equal(np.asarray(a)[keep], np.asarray(b)[keep])


def _has_n(slab, off, ln):
    return np.array([(slab[int(o):int(o) + int(l)] == 0xFF).any() for o, l in zip(off, ln)], dtype=bool)


def _check_decisions(res, ee_raw, ns, ln, has_n, params, exact=True):
    ok, reason, eef = po.decide_batch(
        ee_raw, ns, ln, has_n, thr_kind="maxerrors" if params.maxerrors else "uncert",
        thr=params.maxerrors if params.maxerrors else params.uncert, ambigs=params.ambigs,
        round_flag=params.round, truncate=params.truncate)
    near = res.near_cutoff
    assert np.array_equal(res.accept[~near], ok[~near])
    assert np.array_equal(res.reason[~near], reason[~near])
    return ok, reason, eef


# ---- the reference's own known answers, through the drop-in `bernoulli` module -------------------
def test_kat_through_bernoulli_shim(kat):
    import bernoulli
    assert bernoulli.calculate_errors_PB(kat["testSeq1"], _q1(kat), 0.005) == tuple(kat["pb_expected"])   # test_moira.py:43
    fp = kat["forward_process"]
    ee, ns = bernoulli.calculate_errors_PB(fp["seq"], fp["quals"], 0.005)
    assert ee + ns == fp["ee"]                                                                         # test_moira.py:127
    pp = kat["paired_process"]
    ee, ns = bernoulli.calculate_errors_PB(pp["seq"], pp["quals"], 0.005)
    assert ee + ns == pp["ee"]                                                                         # test_moira.py:128
    assert bernoulli.calculate_errors_PB("", [], 0.005) == (0.0, 0)
    assert bernoulli.calculate_errors_PB("NNnN", [2, 2, 2, 2], 0.005) == (0.0, 4)
    assert bernoulli.calculate_errors_PB("A", [40], 0.005) == (0.0, 0)
    assert bernoulli.calculate_errors_PB("AC", [0, 0], 0.3) == po.pb_c("AC", [1, 1], 0.3)
    with pytest.raises(ValueError):
        bernoulli.calculate_errors_PB("AC", [10, 300], 0.005)


def test_reference_binary_outputs_synthetic(ctx, ref_outputs):
    """4000 seeded reads (N/n, Q=0, L=1..600, six alphas): ee and Ns bit-equal to the unmodified
    reference binary's outputs stored in tests/golden/ref_outputs.npz."""
    slab, off, ln = ref_outputs["syn_slab"], ref_outputs["syn_offsets"], ref_outputs["syn_lengths"]
    alphas = ref_outputs["syn_alpha"]
    for a in np.unique(alphas):
        sel = np.nonzero(alphas == a)[0]
        p = FilterParams(alpha=float(a), exact_ee=True)
        res = ctx.filter_batch(slab, off[sel], ln[sel], p)
        assert not res.numeric.any() and not res.lower_bound.any()
        assert np.array_equal(res.ee, ref_outputs["syn_ee"][sel])
        assert np.array_equal(res.ns, ref_outputs["syn_ns"][sel])
        _check_decisions(res, res.ee, res.ns, ln[sel], _has_n(slab, off[sel], ln[sel]), p)
        assert int(res.counters[L.CNT_READS]) == len(sel)
        assert int(res.counters[L.CNT_ACCEPTED]) == int(res.accept.sum())


def test_forward_fixture_full_pipeline_partition(ctx, forward_records, forward_names, ref_outputs):
    """The reference's forward full-pipeline test (test_moira.py:74-87): fastq -> slab (C parser) ->
    CUDA filter -> collapse -> 122 good / 365 bad uniques with identical member lists."""
    text = gzip.open(os.path.join(ROOT, "tests", "golden", "test1.fastq.gz"), "rb").read()
    slab, off, ln, hoff, hlen, soff, qoff = moira_b200.parse_fastq(text, 33, True)
    p = FilterParams(exact_ee=True, ee_output="final")
    res = ctx.filter_batch(slab, off, ln, p)
    assert np.array_equal(res.ee - res.ns, ref_outputs["forward_ee"])          # raw ee bit-equal to the reference binary
    uniques = {}
    for i, (header, seq, _) in enumerate(forward_records):
        u = uniques.get(seq)
        if u is None:
            uniques[seq] = {"rep": header, "ee": res.ee[i], "names": [header]}
        elif res.ee[i] < u["ee"]:                                               # moira.py:466
            u.update(rep=header, ee=res.ee[i])
            u["names"].insert(0, header)
        else:
            u["names"].append(header)
    good = {u["rep"]: u["names"] for s, u in uniques.items() if u["ee"] <= len(s) * 0.01}
    bad = {u["rep"]: u["names"] for s, u in uniques.items() if not u["ee"] <= len(s) * 0.01}
    assert good == forward_names["good"] and bad == forward_names["bad"]
    # per-read device decision == decision on the representative's own ee
    for i, (_, seq, _) in enumerate(forward_records):
        assert bool(res.accept[i]) == bool(res.ee[i] <= len(seq) * 0.01)


def test_golden_contigs(ctx, contigs, ref_outputs):
    slab, off, ln = moira_b200.pack_reads([c["seq"] for c in contigs], [c["quals"] for c in contigs])
    for exact in (True, False):
        res = ctx.filter_batch(slab, off, ln, FilterParams(exact_ee=exact))
        assert [bool(a) for a in res.accept] == [c["label"] == "good" for c in contigs]
        sel = ~res.lower_bound
        assert np.array_equal(res.ee[sel], ref_outputs["contigs_ee"][sel])
        if exact:
            assert sel.all()


@pytest.mark.parametrize("profile,n,seed", [("v4", 20000, 1), ("v3v4", 6000, 2), ("mixed", 6000, 3), ("ccs", 300, 4)])
def test_profiles_exact_and_decision_modes(ctx, profile, n, seed):
    slab, off, ln = synth.generate(profile, n, seed)
    ee_o, ns_o = po.pb_batch(slab, off, ln, 0.005)
    has_n = _has_n(slab, off, ln)
    p = FilterParams(exact_ee=True)
    res = ctx.filter_batch(slab, off, ln, p)
    assert not res.numeric.any()
    assert np.array_equal(res.ee, ee_o) and np.array_equal(res.ns, ns_o)
    ok, reason, _ = _check_decisions(res, ee_o, ns_o, ln, has_n, p)
    assert np.array_equal((res.flags & L.FLAG_HAS_N) != 0, has_n)
    # decision mode: same partition; exact ee wherever it is not flagged as a lower bound
    pd = FilterParams(exact_ee=False)
    rd = ctx.filter_batch(slab, off, ln, pd)
    assert np.array_equal(rd.accept, res.accept) and np.array_equal(rd.reason, res.reason)
    lb = rd.lower_bound
    assert np.array_equal(rd.ee[~lb], ee_o[~lb])
    assert not rd.accept[lb].any() and np.all(rd.ee[lb] <= ee_o[lb])
    assert int(rd.counters[L.CNT_ACCEPTED]) == int(ok.sum())
    assert int(rd.counters[L.CNT_READS]) == n
    hist = np.bincount(np.minimum(np.floor(res.ee + res.ns), 63).astype(int), minlength=64)
    assert np.array_equal(res.counters[L.CNT_HIST:L.CNT_HIST + 64].astype(np.int64), hist)


@pytest.mark.parametrize("kw", [
    dict(maxerrors=1.0), dict(maxerrors=3.5, round=True), dict(uncert=0.02, ambigs="ignore"),
    dict(ambigs="disallow"), dict(truncate=200), dict(truncate=300, ambigs="disallow", round=True),
    dict(uncert=0.2), dict(maxerrors=60.0), dict(alpha=0.2), dict(alpha=1e-9, uncert=0.05),
])
def test_decision_rules(ctx, kw):
    """A10/A11: truncate, +Ns, floor, maxerrors/uncert, ambigs -- both modes, every rule."""
    slab, off, ln = synth.generate("mixed", 3000, 11)
    for exact in (True, False):
        p = FilterParams(exact_ee=exact, **kw)
        res = ctx.filter_batch(slab, off, ln, p)
        eff = np.minimum(ln, p.truncate) if p.truncate else ln
        ee_o, ns_o = po.pb_batch(slab, off, eff.astype(np.uint32), p.alpha)
        has_n = _has_n(slab, off, eff)
        _check_decisions(res, ee_o, ns_o, ln, has_n, p)
        lb = res.lower_bound
        assert np.array_equal(res.ee[~lb], ee_o[~lb]) and np.array_equal(res.ns, ns_o)
        assert exact is False or not lb.any()
        pf = FilterParams(exact_ee=exact, ee_output="final", **kw)
        rf = ctx.filter_batch(slab, off, ln, pf)
        _, _, eef = po.decide_batch(ee_o, ns_o, ln, has_n, thr_kind="maxerrors" if p.maxerrors else "uncert",
                                    thr=p.maxerrors or p.uncert, ambigs=p.ambigs, round_flag=p.round, truncate=p.truncate)
        assert np.array_equal(rf.ee[~lb], eef[~lb])


def test_escalation_ladder_every_rung(ctx):
    """Reads needing 1 .. >1024 PMF entries: thread-, warp- and block-per-read rungs all bit-exact."""
    rng = np.random.default_rng(5)
    seqs, quals = [], []
    for L_, qlo, qhi in ((300, 30, 41), (300, 10, 20), (400, 3, 9), (600, 1, 4), (1500, 0, 3), (1500, 2, 6),
                         (2600, 0, 2), (3000, 1, 3), (40, 0, 2), (5000, 25, 40)):
        for _ in range(6):
            seqs.append("".join(rng.choice(list("ACGTN"), size=L_, p=[.245, .245, .25, .25, .01])))
            quals.append([int(v) for v in rng.integers(qlo, qhi, size=L_)])
    slab, off, ln = moira_b200.pack_reads(seqs, quals)
    ee_o, ns_o = po.pb_batch(slab, off, ln, 0.005)
    assert ee_o.max() > 1100 and ee_o.min() < 2
    res = ctx.filter_batch(slab, off, ln, FilterParams(exact_ee=True))
    assert not res.numeric.any() and not res.lower_bound.any()
    assert np.array_equal(res.ee, ee_o) and np.array_equal(res.ns, ns_o)
    # decision mode with a cutoff far above the first-pass K (maxerrors 700 -> needs the ladder too)
    p = FilterParams(exact_ee=False, maxerrors=700.0)
    rd = ctx.filter_batch(slab, off, ln, p)
    ok, _, _ = po.decide_batch(ee_o, ns_o, ln, _has_n(slab, off, ln), thr_kind="maxerrors", thr=700.0,
                               ambigs="treat_as_errors", round_flag=False, truncate=None)
    assert np.array_equal(rd.accept, ok)


def test_poisson_and_expected_error_modes(ctx, kat):
    slab, off, ln = synth.generate("mixed", 1500, 21)
    ee_o, ns_o, lam_o = po.poisson_batch(slab, off, ln, 0.005)
    has_n = _has_n(slab, off, ln)
    pe = FilterParams(error_calc="expected_error", exact_ee=True)
    re_ = ctx.filter_batch(slab, off, ln, pe)
    assert np.array_equal(re_.ee, lam_o) and np.array_equal(re_.ns, ns_o)      # sequential sum: bit-exact
    _check_decisions(re_, lam_o, ns_o, ln, has_n, pe)
    pp = FilterParams(error_calc="poisson", exact_ee=True)
    rp = ctx.filter_batch(slab, off, ln, pp)
    assert not rp.numeric.any()
    assert np.allclose(rp.ee, ee_o, rtol=1e-12, atol=1e-13)                    # device exp/pow vs glibc: stated tolerance
    ok, _, _ = _check_decisions(rp, ee_o, ns_o, ln, has_n, pp)
    rd = ctx.filter_batch(slab, off, ln, FilterParams(error_calc="poisson", exact_ee=False))
    near = rd.near_cutoff | rp.near_cutoff
    assert np.array_equal(rd.accept[~near], ok[~near])
    # the reference's own Poisson KAT (test_moira.py:44-45), to the stated tolerance
    s, o, l = moira_b200.pack_reads([kat["testSeq1"]], [_q1(kat)], lower_n_ambiguous=False)
    r1 = ctx.filter_batch(s, o, l, pp)
    assert abs(r1.ee[0] - kat["poisson_expected"][0]) <= 1e-12 * kat["poisson_expected"][0]


def test_edge_batches(ctx):
    p = FilterParams()
    r0 = ctx.filter_batch(np.zeros(16, np.uint8), np.zeros(0, np.uint64), np.zeros(0, np.uint32), p)
    assert r0.ee.size == 0 and int(r0.counters[L.CNT_READS]) == 0
    # ragged: empty read, 1 base, 15/16/17 bases, all-N, garbage in the row padding
    seqs = ["", "A", "A" * 15, "C" * 16, "G" * 17, "N" * 33, "ACGT" * 70]
    quals = [[], [2], [30] * 15, [3] * 16, [40] * 17, [2] * 33, [38] * 280]
    slab, off, ln = moira_b200.pack_reads(seqs, quals)
    slab = slab.copy()
    for o, l in zip(off, ln):
        pad = (int(l) + 15) // 16 * 16 - int(l)
        slab[int(o) + int(l):int(o) + int(l) + pad] = np.random.default_rng(1).integers(0, 256, pad)
    ee_o = np.array([po.pb_c(s, q, 0.005)[0] for s, q in zip(seqs, quals)])
    res = ctx.filter_batch(slab, off, ln, p)
    assert np.array_equal(res.ee, ee_o)
    assert list(res.ns) == [0, 0, 0, 0, 0, 33, 0]
    with pytest.raises(moira_b200.MoiraError) as ei:
        ctx.filter_batch(slab, off, ln, FilterParams(alpha=1.5))
    assert ei.value.code == L.ERR_BAD_ALPHA
    with pytest.raises(moira_b200.MoiraError):
        ctx.filter_batch(slab, off + 3, ln, p)


def test_device_resident_api_and_properties_at_scale(ctx):
    """BASELINE config C2 at its full size (10 M x 253 bp), device-resident: decision mode == exact mode,
    permutation invariance, counters == sums over flags, and a sampled oracle check (both ladder
    sub-batches)."""
    import torch
    n = 10_000_000
    dev = torch.device("cuda", 0)
    slab = synth.generate_v4_device(n, 1234, dev)
    ee = torch.empty(n, dtype=torch.float64, device=dev)
    ns = torch.empty(n, dtype=torch.int32, device=dev)
    fl = torch.empty(n, dtype=torch.uint8, device=dev)
    cnt = torch.zeros(L.N_COUNTERS, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    outs = {}
    for exact in (False, True):
        cnt.zero_()
        ctx.filter_device(slab.data_ptr(), None, None, synth.V4_STRIDE, synth.V4_LEN, n,
                          FilterParams(exact_ee=exact), ee.data_ptr(), ns.data_ptr(), fl.data_ptr(),
                          cnt.data_ptr(), stream)
        torch.cuda.synchronize()
        outs[exact] = (ee.cpu().numpy().copy(), ns.cpu().numpy().copy(), fl.cpu().numpy().copy(), cnt.cpu().numpy().copy())
    (ee_d, ns_d, fl_d, c_d), (ee_x, ns_x, fl_x, c_x) = outs[False], outs[True]
    assert np.array_equal(fl_d & 1, fl_x & 1) and np.array_equal(ns_d, ns_x)
    lb = (fl_d & L.FLAG_LOWER_BOUND) != 0
    assert np.array_equal(ee_d[~lb], ee_x[~lb]) and not ((fl_x & L.FLAG_LOWER_BOUND) != 0).any()
    assert c_x[L.CNT_READS] == n and c_x[L.CNT_ACCEPTED] == int((fl_x & 1).sum()) == c_d[L.CNT_ACCEPTED]
    assert 0.7 < c_x[L.CNT_ACCEPTED] / n < 0.9                      # ~80 % accepted
    # sampled bit-exact check against the oracle
    idx = np.concatenate([np.random.default_rng(9).choice(n, 3000, replace=False), np.arange(n - 1000, n)])
    rows = slab[torch.as_tensor(idx, device=dev)].cpu().numpy()
    off = np.arange(len(idx), dtype=np.uint64) * synth.V4_STRIDE
    ee_o, ns_o = po.pb_batch(rows.reshape(-1), off, np.full(len(idx), synth.V4_LEN, np.uint32), 0.005)
    assert np.array_equal(ee_x[idx], ee_o) and np.array_equal(ns_x[idx], ns_o)
    # permutation invariance (reads are independent, moira.py:455-487)
    perm = torch.randperm(n, device=dev)
    slab_p = slab[perm].contiguous()
    cnt.zero_()
    ctx.filter_device(slab_p.data_ptr(), None, None, synth.V4_STRIDE, synth.V4_LEN, n, FilterParams(exact_ee=True),
                      ee.data_ptr(), ns.data_ptr(), fl.data_ptr(), cnt.data_ptr(), stream)
    torch.cuda.synchronize()
    assert np.array_equal(ee.cpu().numpy(), ee_x[perm.cpu().numpy()])
    assert np.array_equal(cnt.cpu().numpy(), c_x)


def test_async_submit_wait_pinned(ctx):
    slab, off, ln = synth.generate("v4", 50000, 77)
    bufs, outs, tickets = [], [], []
    p = FilterParams(exact_ee=False)
    for k in range(3):
        pb = moira_b200.PinnedBuffer(slab.nbytes)
        pb.u8[:] = slab
        out = moira_b200.FilterResult(np.empty(len(ln)), np.empty(len(ln), np.int32), np.empty(len(ln), np.uint8),
                                      np.zeros(L.N_COUNTERS, np.uint64))
        tickets.append(ctx.submit(pb.u8, off, ln, p, out))
        bufs.append(pb)
        outs.append(out)
    for t in tickets:
        ctx.wait(t)
    ref = ctx.filter_batch(slab, off, ln, p)
    for out in outs:
        assert np.array_equal(out.ee, ref.ee) and np.array_equal(out.flags, ref.flags)
        assert _same_counters(out.counters, ref.counters)
    for pb in bufs:
        pb.free()


def test_streaming_fastq_entry_point(ctx):
    """moira_filter_fastq (parse ranges -> pinned slabs -> async submits) == parse_fastq + filter_batch,
    across a range boundary (> 64 MB of text) and with a trailing partial record."""
    text = gzip.open(os.path.join(ROOT, "tests", "golden", "test1.fastq.gz"), "rb").read()
    p = FilterParams(exact_ee=True, ee_output="final")
    slab, off, ln, *_ = moira_b200.parse_fastq(text, 33, True)
    ref = ctx.filter_batch(slab, off, ln, p)
    res, lengths = ctx.filter_fastq(text, p)
    assert np.array_equal(lengths, ln) and np.array_equal(res.ee, ref.ee) and np.array_equal(res.flags, ref.flags)
    assert _same_counters(res.counters, ref.counters)
    big = text * 130 + b"@partial\nACGT\n"                      # ~74 MB: two streaming ranges
    assert len(big) > (64 << 20)
    res2, lengths2 = ctx.filter_fastq(big, p)
    assert len(lengths2) == 130 * len(ln)
    assert np.array_equal(lengths2, np.tile(ln, 130)) and np.array_equal(res2.ee, np.tile(ref.ee, 130))
    assert np.array_equal(res2.flags, np.tile(ref.flags, 130))
    assert _same_counters(res2.counters, ref.counters * np.uint64(130))
    with pytest.raises(moira_b200.MoiraError) as ei:
        ctx.filter_fastq(text[:5000] + b"@bad\nACGT\n+\nIII\n" + text[5000:], p)
    assert ei.value.code == L.ERR_PARSE


def test_length_bucketing_of_ragged_batches(ctx):
    """>= 32768 ragged reads take the on-device counting sort by length (per-bucket K); results must be
    identical to the unsorted path and to the oracle, in input order."""
    slab, off, ln = synth.generate("mixed", 40000, 31)
    ee_o, ns_o = po.pb_batch(slab, off, ln, 0.005)
    has_n = _has_n(slab, off, ln)
    for kw in (dict(), dict(maxerrors=3.0), dict(truncate=333), dict(uncert=0.08)):
        for exact in (True, False):
            p_on = FilterParams(exact_ee=exact, length_sort=1, **kw)
            p_off = FilterParams(exact_ee=exact, length_sort=2, **kw)
            r_on = ctx.filter_batch(slab, off, ln, p_on)
            r_off = ctx.filter_batch(slab, off, ln, p_off)
            assert np.array_equal(r_on.accept, r_off.accept) and np.array_equal(r_on.ns, r_off.ns)
            assert np.array_equal(r_on.reason, r_off.reason)
            both = ~r_on.lower_bound & ~r_off.lower_bound
            assert np.array_equal(r_on.ee[both], r_off.ee[both])
            assert int(r_on.counters[L.CNT_READS]) == 40000
            assert np.array_equal(r_on.counters[:5], r_off.counters[:5])
            if not kw:
                assert np.array_equal(r_on.ns, ns_o)
                lb = r_on.lower_bound
                assert np.array_equal(r_on.ee[~lb], ee_o[~lb]) and (exact is False or not lb.any())
                _check_decisions(r_on, ee_o, ns_o, ln, has_n, p_on)
    for calc in ("poisson", "expected_error"):
        r_on = ctx.filter_batch(slab, off, ln, FilterParams(error_calc=calc, length_sort=1))
        r_off = ctx.filter_batch(slab, off, ln, FilterParams(error_calc=calc, length_sort=2))
        assert np.array_equal(r_on.ee, r_off.ee) and np.array_equal(r_on.flags, r_off.flags)


def test_near_cutoff_band_is_flagged_and_counted(ctx):
    """Reads whose statistic lies within 1e-12 (relative) of the cutoff carry MOIRA_FLAG_NEAR_CUTOFF and
    are counted (north_star's tolerance band); outside the band the flag is clear."""
    slab, off, ln = synth.generate("v4", 64, 5)
    ee_o, ns_o = po.pb_batch(slab, off, ln, 0.005)
    i = int(np.argmax((ee_o > 0.5) & (ns_o == 0)))
    target = float(ee_o[i])
    for thr, expect_near in ((target, True), (target * (1 - 3e-13), True), (target * (1 + 1e-9), False)):
        r = ctx.filter_batch(slab, off, ln, FilterParams(maxerrors=thr, ambigs="ignore"))
        assert bool(r.near_cutoff[i]) is expect_near
        assert bool(r.accept[i]) == (target <= thr)
        assert int(r.counters[L.CNT_NEAR_CUTOFF]) == int(r.near_cutoff.sum())


def test_q6_transport_format_gives_identical_results(ctx):
    """Host slabs sent as 6-bit images (3/4 of the PCIe bytes) are expanded on the device: same outputs."""
    for profile, n in (("v4", 30000), ("mixed", 5000)):
        slab, off, ln = synth.generate(profile, n, 17)
        img = moira_b200.pack_q6(slab)
        for exact in (False, True):
            r8 = ctx.filter_batch(slab, off, ln, FilterParams(exact_ee=exact))
            r6 = ctx.filter_batch(img, off, ln, FilterParams(exact_ee=exact, slab_format="q6"))
            assert np.array_equal(r8.ee, r6.ee) and np.array_equal(r8.ns, r6.ns) and np.array_equal(r8.flags, r6.flags)
            assert _same_counters(r8.counters, r6.counters)


@pytest.mark.gpu
def test_submit_failing_midway_leaves_the_context_usable(ctx):
    """A batch whose later chunk is malformed fails with BAD_ARG after earlier chunks were enqueued; the
    streams are drained before the error returns and the next call gives the same results as ever."""
    slab, off, ln = synth.generate("v4", 400000, 5)             # ~100 MB: several 32 MB chunks
    p = FilterParams(exact_ee=False)
    good = ctx.filter_batch(slab, off, ln, p)
    bad_off = off.copy()
    bad_off[-10] += 8                                           # not a multiple of 16, in the last chunk
    out = moira_b200.FilterResult(np.empty(len(ln)), np.empty(len(ln), np.int32), np.empty(len(ln), np.uint8),
                                  np.zeros(L.N_COUNTERS, np.uint64))
    with pytest.raises(moira_b200.MoiraError) as ei:
        ctx.submit(slab, bad_off, ln, p, out)
    assert ei.value.code == L.ERR_BAD_ARG and "multiple of 16" in ei.value.message
    for _ in range(L.MAX_INFLIGHT + 1):                         # no ticket leaked by the failed submission
        again = ctx.filter_batch(slab, off, ln, p)
        assert np.array_equal(again.flags, good.flags) and _same_counters(again.counters, good.counters)


@pytest.mark.gpu
def test_device_fastq_parser_agrees_with_host_parser(ctx, monkeypatch):
    """moira_filter_fastq parses on the device by default.  Ragged reads, N / n, '@' and '+' as quality
    characters, CRLF and padded lines, an unterminated last record and a trailing partial record must give what
    the host parser (pinned on the reference's records in test_abi.py) plus filter_batch give; errors are the
    host parser's."""
    rng = np.random.default_rng(21)
    recs = []
    for i in range(40000):
        n = int(rng.integers(1, 400)) if i % 4 else 253
        seq = "".join(rng.choice(list("ACGTNn"), size=n, p=[.24, .24, .24, .24, .03, .01]))
        q = rng.integers(33, 74, size=n).astype(np.uint8)
        if i % 3 == 0:
            q[0] = ord("@")
        if i % 5 == 0:
            q[-1] = ord("+")
        eol = "\r\n" if i % 7 == 0 else "\n"
        pad = "  " if i % 11 == 0 else ""
        recs.append("@r%d some text%s%s%s%s%s+%s%s%s%s" % (i, eol, pad, seq, pad, eol, eol, q.tobytes().decode(), "\t" if i % 13 == 0 else "", eol))
    body = "".join(recs).encode()
    p = FilterParams(exact_ee=True, ee_output="final")
    for tail in (b"", b"@last\nACGTN\n+\nIIII#", b"@partial\nACGT\n+\n"):
        text = body + tail
        slab, off, ln, *_ = moira_b200.parse_fastq(text, 33, True)
        ref = ctx.filter_batch(slab, off, ln, p)
        res, lengths = ctx.filter_fastq(text, p)
        assert len(lengths) == len(ln) == 40000 + (1 if tail.startswith(b"@last") else 0)
        assert np.array_equal(lengths, ln) and np.array_equal(res.ee, ref.ee) and np.array_equal(res.ns, ref.ns)
        assert np.array_equal(res.flags, ref.flags) and _same_counters(res.counters, ref.counters)
    # the host-parse variant of the same entry point (a context created with MOIRA_B200_HOST_PARSE=1)
    monkeypatch.setenv("MOIRA_B200_HOST_PARSE", "1")
    c2 = moira_b200.Context(0)
    try:
        res_h, lengths_h = c2.filter_fastq(text, p)
        assert np.array_equal(lengths_h, ln) and np.array_equal(res_h.ee, ref.ee) and np.array_equal(res_h.flags, ref.flags)
    finally:
        c2.close()
        monkeypatch.delenv("MOIRA_B200_HOST_PARSE")
    # several 64 MB chunks: the cuts are guessed from the text's local structure and verified by the device's line count
    big = body * 5 + b"@last\nACGTN\n+\nIIII#"
    slab, off, ln, *_ = moira_b200.parse_fastq(big, 33, True)
    ref_big = ctx.filter_batch(slab, off, ln, p)
    assert len(big) > 70 * 2**20 and len(ln) == 200001
    res_big, lengths_big = ctx.filter_fastq(big, p)
    assert np.array_equal(lengths_big, ln) and np.array_equal(res_big.ee, ref_big.ee) and np.array_equal(res_big.flags, ref_big.flags)
    assert _same_counters(res_big.counters, ref_big.counters)
    monkeypatch.setenv("MOIRA_B200_FQ_COUNT", "1")            # the same with every cut counted on the host
    c3 = moira_b200.Context(0)
    try:
        res_c, lengths_c = c3.filter_fastq(big, p)
        assert np.array_equal(lengths_c, ln) and np.array_equal(res_c.ee, ref_big.ee) and np.array_equal(res_c.flags, ref_big.flags)
    finally:
        c3.close()
        monkeypatch.delenv("MOIRA_B200_FQ_COUNT")
    # lines shorter than the index estimate (8 bytes): the device reports the overflow, the call counts and runs again
    tiny = b"".join(b"@%d\n%s\n+\n%s\n" % (i, b"ACGTN"[i % 5:i % 5 + 1], b"I5#"[i % 3:i % 3 + 1]) for i in range(50000))
    slab, off, ln, *_ = moira_b200.parse_fastq(tiny, 33, True)
    ref_t = ctx.filter_batch(slab, off, ln, p)
    res_t, lengths_t = ctx.filter_fastq(tiny, p)
    assert np.array_equal(lengths_t, ln) and np.array_equal(res_t.ee, ref_t.ee) and np.array_equal(res_t.flags, ref_t.flags)
    # errors: detected on the device, raised with the host parser's message
    good = "".join(recs[:3000]).encode()
    for bad, word in ((b"@bad\nACGT\n+\nIII\n", "LengthMismatchError"), (b"@bad\n\n+\nIII\n", "EmptySeqError"),
                      (b"@bad\nACG\n+\n\n", "EmptyQualError")):
        with pytest.raises(moira_b200.MoiraError) as ei:
            ctx.filter_fastq(good + bad + good, p)
        assert ei.value.code == L.ERR_PARSE and word in ei.value.message and "3000" in ei.value.message
    with pytest.raises(moira_b200.MoiraError) as ei:
        ctx.filter_fastq(good, p, fastq_offset=-190)
    assert ei.value.code == L.ERR_BAD_QUALITY
    # the context is fine afterwards
    res3, _ = ctx.filter_fastq(text, p)
    assert np.array_equal(res3.ee, ref.ee)


@pytest.mark.parametrize("profile,n,kw", [
    ("v4", 40000, dict()), ("v4", 40000, dict(maxerrors=1.5)), ("v4", 40000, dict(uncert=0.02, ambigs="ignore")),
    ("v3v4", 20000, dict()), ("v3v4", 20000, dict(maxerrors=6.0, round=True)), ("mixed", 20000, dict(maxerrors=4.2, length_sort=2)),
    ("mixed", 20000, dict(truncate=250, ambigs="disallow", length_sort=2)), ("v4", 40000, dict(alpha=0.2)),
    ("v4", 40000, dict(alpha=1e-6, maxerrors=2.0)),
])
def test_cascaded_first_pass_gives_identical_results(ctx, profile, n, kw):
    """Decisions that need 3..8 PMF entries: the two-entry sweep + Newton-bound rejects + second sweep (cascade = 1)
    must write what the single full-K sweep (cascade = 2) writes -- ee, Ns, flags, counters, bit for bit -- in both
    modes, and both must agree with the oracle."""
    slab, off, ln = synth.generate(profile, n, 31)
    for exact in (False, True):
        one = ctx.filter_batch(slab, off, ln, FilterParams(exact_ee=exact, cascade=2, **kw))
        two = ctx.filter_batch(slab, off, ln, FilterParams(exact_ee=exact, cascade=1, **kw))
        auto = ctx.filter_batch(slab, off, ln, FilterParams(exact_ee=exact, **kw))
        for r in (two, auto):
            assert np.array_equal(r.ee, one.ee) and np.array_equal(r.ns, one.ns) and np.array_equal(r.flags, one.flags)
            same = (np.arange(L.N_COUNTERS) != L.CNT_ESCALATED) & (np.arange(L.N_COUNTERS) != L.CNT_FP64_OPS) & (np.arange(L.N_COUNTERS) != L.CNT_CLASSIFIED)   # the counters that describe the route, not the result
            assert np.array_equal(r.counters[same], one.counters[same])
        p = FilterParams(exact_ee=exact, **{k: v for k, v in kw.items() if k != "length_sort"})
        eff = np.minimum(ln, p.truncate) if p.truncate else ln
        ee_o, ns_o = po.pb_batch(slab, off, eff.astype(np.uint32), p.alpha)
        _check_decisions(two, ee_o, ns_o, ln, _has_n(slab, off, eff), p)
        lb = two.lower_bound
        assert np.array_equal(two.ee[~lb], ee_o[~lb]) and np.all(two.ee[lb] <= ee_o[lb]) and not (exact and lb.any())


def test_cascade_pilot_chooses_per_batch(ctx):
    """Batches large enough for the pilot launch (>= 8 x 2 tiles per warp): on clean-or-hopeless reads the pilot keeps
    the cascade, on reads that mostly need 3-4 entries it switches the rest of the batch to the full-K sweep; either
    way the results equal the single-sweep ones."""
    import torch
    dev = torch.device("cuda", 0)
    n = 1_400_000
    g = torch.Generator(device=dev)
    g.manual_seed(5)
    borderline = torch.randint(26, 29, (n, synth.V4_STRIDE), dtype=torch.uint8, device=dev, generator=g)   # mean error count ~0.5: j* is 2 or 3
    borderline[:, synth.V4_LEN:] = 0xFD
    slabs = {"v4": synth.generate_v4_device(n, 77, dev), "borderline": borderline}
    stream = torch.cuda.current_stream().cuda_stream
    for name, slab in slabs.items():
        outs = []
        for exact, cascade in ((False, 2), (False, 0), (True, 2), (True, 0)):
            ee = torch.empty(n, dtype=torch.float64, device=dev)
            ns = torch.empty(n, dtype=torch.int32, device=dev)
            fl = torch.empty(n, dtype=torch.uint8, device=dev)
            cnt = torch.zeros(L.N_COUNTERS, dtype=torch.int64, device=dev)
            before = ctx.launch_count
            ctx.filter_device(slab.data_ptr(), None, None, synth.V4_STRIDE, synth.V4_LEN, n,
                              FilterParams(exact_ee=exact, cascade=cascade), ee.data_ptr(), ns.data_ptr(), fl.data_ptr(),
                              cnt.data_ptr(), stream)
            torch.cuda.synchronize()
            outs.append((ee.cpu().numpy(), ns.cpu().numpy(), fl.cpu().numpy(), cnt.cpu().numpy(), ctx.launch_count - before))
        same = (np.arange(L.N_COUNTERS) != L.CNT_ESCALATED) & (np.arange(L.N_COUNTERS) != L.CNT_FP64_OPS) & (np.arange(L.N_COUNTERS) != L.CNT_CLASSIFIED)
        for a, b in ((0, 1), (2, 3)):
            for k in range(3):
                assert np.array_equal(outs[a][k], outs[b][k]), (name, a, k)
            assert np.array_equal(outs[a][3][same], outs[b][3][same]), (name, a)
        esc = outs[1][3][L.CNT_ESCALATED] / n
        assert (esc < 0.3) if name == "v4" else (0.05 < esc < 0.12)   # borderline batch: only the pilot's reads took the two-entry sweep
        assert outs[0][4] == 1 and outs[1][4] == 5      # one sweep | pilot, verdict, two candidates, second sweep
        idx = np.random.default_rng(3).choice(n, 2000, replace=False)
        rows = slab[torch.as_tensor(idx, device=dev)].cpu().numpy()
        off = np.arange(len(idx), dtype=np.uint64) * synth.V4_STRIDE
        ee_o, ns_o = po.pb_batch(rows.reshape(-1), off, np.full(len(idx), synth.V4_LEN, np.uint32), 0.005)
        assert np.array_equal(outs[3][0][idx], ee_o) and np.array_equal(outs[3][1][idx], ns_o)


def test_pilotless_cascade_learns_from_finished_batches():
    """Batches too small for the pilot take the cascade blindly; when a finished batch's counters show that most reads
    were escalated, the following batches sweep once with all entries -- same results either way."""
    rng = np.random.default_rng(8)
    n = 60000
    slab = np.full((n, 256), 0xFD, np.uint8)
    slab[:, :253] = rng.integers(26, 29, (n, 253), dtype=np.uint8)          # mean error count ~0.5: j* is 2 or 3 for most reads
    off = np.arange(n, dtype=np.uint64) * 256
    ln = np.full(n, 253, np.uint32)
    c = moira_b200.Context(0)
    try:
        p = FilterParams(exact_ee=False)
        first = c.filter_batch(slab.reshape(-1), off, ln, p)
        again = c.filter_batch(slab.reshape(-1), off, ln, p)
        assert int(first.counters[L.CNT_ESCALATED]) > 0.35 * n and int(again.counters[L.CNT_ESCALATED]) == 0
        assert np.array_equal(first.ee, again.ee) and np.array_equal(first.flags, again.flags) and _same_counters(first.counters, again.counters)
        ee_o, ns_o = po.pb_batch(slab.reshape(-1), off, ln, 0.005)
        lb = again.lower_bound
        assert np.array_equal(again.ee[~lb], ee_o[~lb]) and np.array_equal(again.accept, (ee_o + ns_o) <= 253 * 0.01)
    finally:
        c.close()
