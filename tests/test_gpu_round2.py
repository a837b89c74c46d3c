"""GPU (-m gpu), round 2: row marks from the slab's producers, the N-context path (shards of one slab on several
contexts == one context), the counters' all-reduce inside the library, the executed-FP64 counter, device-resident
ragged batches with host-known lengths, and a 20 000-read CCS (1 500 bp) parity sample."""
import numpy as np
import pytest

import moira_b200
from moira_b200 import FilterParams, synth
from moira_b200 import _lib as L
from oracle import py_oracle as po

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _dev_arrays(n):
    dev = torch.device("cuda", 0)
    return (torch.empty(n, dtype=torch.float64, device=dev), torch.empty(n, dtype=torch.int32, device=dev),
            torch.empty(n, dtype=torch.uint8, device=dev), torch.zeros(L.N_COUNTERS, dtype=torch.int64, device=dev))


def _marks_numpy(slab, off, ln, truncate=0):
    out = np.zeros(len(ln), np.uint32)
    for r, (o, l) in enumerate(zip(off, ln)):
        eff = min(int(l), truncate) if truncate else int(l)
        row = slab[int(o):int(o) + eff]
        out[r] = int((row >= 0xFE).sum()) | ((1 << 31) if (row == 0xFF).any() else 0)
    return out


def _same_but_diagnostics(a, b):
    keep = np.ones(L.N_COUNTERS, bool)
    keep[[L.CNT_ESCALATED, L.CNT_FP64_OPS, L.CNT_CLASSIFIED]] = False
    return np.array_equal(np.asarray(a)[keep], np.asarray(b)[keep])


@pytest.mark.parametrize("profile,n,seed,truncate", [("v4", 40000, 11, 0), ("v4", 40000, 12, 200), ("mixed", 40000, 13, 0),
                                                      ("mixed", 40000, 14, 250), ("v3v4", 40000, 15, 0)])
def test_row_marks_producer_and_filter(ctx, profile, n, seed, truncate):
    """moira_count_marks_device == a numpy count; a filter call that is GIVEN the marks writes exactly what the
    self-counting call writes (every array, every counter) in decision and exact mode."""
    slab, off, ln = synth.generate(profile, n, seed)
    slab = slab.copy()
    rng = np.random.default_rng(seed)
    pos = rng.integers(0, len(slab), 2000)
    slab[pos] = np.where(slab[pos] == 0xFD, 0xFD, 0xFE)          # some lower-case n as well
    stride = int(off[1] - off[0])
    dev = torch.device("cuda", 0)
    d_slab = torch.from_numpy(slab).to(dev)
    uniform = bool((ln == ln[0]).all())
    d_len = None if uniform else torch.from_numpy(ln.astype(np.int32)).to(dev)
    marks = torch.zeros(n, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    ctx.count_marks_device(d_slab.data_ptr(), None, None if uniform else d_len.data_ptr(), stride, int(ln[0]) if uniform else 0, n,
                           marks.data_ptr(), truncate, stream)
    torch.cuda.synchronize()
    assert np.array_equal(marks.cpu().numpy().view(np.uint32), _marks_numpy(slab, off, ln, truncate))
    for exact in (False, True):
        got = []
        for use in (False, True):
            ee, ns, fl, cnt = _dev_arrays(n)
            p = FilterParams(exact_ee=exact, truncate=truncate or None, max_length=int(ln.max()), min_length=int(ln.min()))
            ctx.filter_device(d_slab.data_ptr(), None, None if uniform else d_len.data_ptr(), stride, int(ln[0]) if uniform else 0, n, p,
                              ee.data_ptr(), ns.data_ptr(), fl.data_ptr(), cnt.data_ptr(), stream, marks.data_ptr() if use else None)
            torch.cuda.synchronize()
            got.append((ee.cpu().numpy(), ns.cpu().numpy(), fl.cpu().numpy(), cnt.cpu().numpy()))
        assert np.array_equal(got[0][0], got[1][0]) and np.array_equal(got[0][1], got[1][1]) and np.array_equal(got[0][2], got[1][2])
        assert _same_but_diagnostics(got[0][3], got[1][3])
        assert got[1][3][L.CNT_FP64_OPS] > 0
    # ... and the exact results are the oracle's
    eff = np.minimum(ln, truncate) if truncate else ln
    ee_o, ns_o = po.pb_batch(slab, off, eff.astype(np.uint32), 0.005)
    assert np.array_equal(got[1][0], ee_o) and np.array_equal(got[1][1], ns_o)


def test_q6_transport_leaves_row_marks(ctx):
    """The 6-bit expansion kernel produces the row marks of uniform rows; results equal the byte-slab path's."""
    slab, off, ln = synth.generate("v4", 60000, 21)
    img = moira_b200.pack_q6(slab)
    for exact in (False, True):
        for trunc in (None, 100):
            a = ctx.filter_batch(slab, off, ln, FilterParams(exact_ee=exact, truncate=trunc))
            b = ctx.filter_batch(img, off, ln, FilterParams(exact_ee=exact, truncate=trunc, slab_format="q6"))
            assert np.array_equal(a.ee, b.ee) and np.array_equal(a.ns, b.ns) and np.array_equal(a.flags, b.flags)
            assert _same_but_diagnostics(a.counters, b.counters)
    ee_o, ns_o = po.pb_batch(slab, off, ln, 0.005)
    b = ctx.filter_batch(img, off, ln, FilterParams(exact_ee=True, slab_format="q6"))
    assert np.array_equal(b.ee, ee_o) and np.array_equal(b.ns, ns_o)


def test_shards_on_several_contexts_equal_one_context(ctx):
    """The N-rank path on the GPU: contiguous shards of one slab on separate contexts (own streams, workspaces, queues),
    concatenated in rank order, are the single-context result; the shard counters add up to its counters."""
    from moira_b200.shard import shard_slab
    slab, off, ln = synth.generate("mixed", 90000, 31)
    for exact in (False, True):
        p = FilterParams(exact_ee=exact)
        whole = ctx.filter_batch(slab, off, ln, p)
        for world in (2, 3):
            ctxs = [moira_b200.Context(0) for _ in range(world)]
            try:
                parts, tickets = [], []
                for r, c in enumerate(ctxs):
                    b, e, byte_b, byte_e = shard_slab(off, ln, r, world)
                    sub_off = (off[b:e] - np.uint64(byte_b)).astype(np.uint64)
                    out = moira_b200.FilterResult(np.empty(e - b), np.empty(e - b, np.int32), np.empty(e - b, np.uint8),
                                                  np.zeros(L.N_COUNTERS, np.uint64))
                    sub_slab = np.ascontiguousarray(slab[byte_b:byte_e])
                    sub_len = np.ascontiguousarray(ln[b:e])
                    tickets.append((c, c.submit(sub_slab, sub_off, sub_len, p, out), (sub_slab, sub_off, sub_len)))   # all in flight at once
                    parts.append(out)
                for c, t, _keep in tickets:
                    c.wait(t)
                ee, fl = np.concatenate([q.ee for q in parts]), np.concatenate([q.flags for q in parts])
                total = np.sum([q.counters for q in parts], axis=0)
                assert np.array_equal(np.concatenate([q.ns for q in parts]), whole.ns)
                if exact:
                    assert np.array_equal(ee, whole.ee) and np.array_equal(fl, whole.flags)
                    assert _same_but_diagnostics(total, whole.counters)
                else:
                    # decision mode: WHICH certain rejects carry a lower bound instead of the exact statistic depends on the
                    # first-pass K of the batch a read travels in; decisions, reasons and every exact value are the same
                    canon = L.FLAG_ACCEPT | L.FLAG_REASON_MASK | L.FLAG_HAS_N
                    assert np.array_equal(fl & canon, whole.flags & canon)
                    both = ((fl | whole.flags) & L.FLAG_LOWER_BOUND) == 0
                    assert np.array_equal(ee[both], whole.ee[both]) and both.mean() > 0.5
                    assert np.array_equal(total[:L.CNT_NEAR_CUTOFF], whole.counters[:L.CNT_NEAR_CUTOFF])
                assert int(total[L.CNT_ACCEPTED]) == int(sum((q.flags & 1).sum() for q in parts))
            finally:
                for c in ctxs:
                    c.close()


def test_counters_allreduce_inside_the_library(ctx):
    """moira_comm_* + moira_reduce_counters*: a communicator of one rank on this box's GPU 0 (sum == identity), host and
    device variants; with two GPUs, two contexts in one process (moira_comm_init_all / moira_reduce_counters_all)."""
    c = moira_b200.Context(0)
    try:
        assert c.comm_info() == (0, 1)
        with pytest.raises(moira_b200.MoiraError):
            c.reduce_counters(np.zeros(L.N_COUNTERS, np.uint64))               # no communicator yet
        c.comm_init(moira_b200.comm_unique_id(), 0, 1)
        assert c.comm_info() == (0, 1)
        v = np.arange(L.N_COUNTERS, dtype=np.uint64) * np.uint64(3) + np.uint64(1 << 40)
        assert np.array_equal(c.reduce_counters(v.copy()), v)
        d = torch.from_numpy(v.astype(np.int64)).to("cuda:0")
        c.reduce_counters_device(d.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert np.array_equal(d.cpu().numpy().astype(np.uint64), v)
    finally:
        c.close()
    if torch.cuda.device_count() >= 2:
        ctxs = [moira_b200.Context(0), moira_b200.Context(1)]
        try:
            moira_b200.comm_init_all(ctxs)
            assert [x.comm_info() for x in ctxs] == [(0, 2), (1, 2)]
            slab, off, ln = synth.generate("v4", 50000, 41)
            half = 25000
            outs = [ctxs[0].filter_batch(slab[:half * 256], off[:half], ln[:half], FilterParams(exact_ee=False)),
                    ctxs[1].filter_batch(slab[half * 256:], off[half:] - off[half], ln[half:], FilterParams(exact_ee=False))]
            want = outs[0].counters + outs[1].counters
            red = moira_b200.reduce_counters_all(ctxs, [o.counters.copy() for o in outs])
            assert np.array_equal(red[0], want) and np.array_equal(red[1], want)
            assert int(want[L.CNT_ACCEPTED]) == int((outs[0].flags & 1).sum() + (outs[1].flags & 1).sum())
        finally:
            for x in ctxs:
                x.close()


def test_device_ragged_batch_with_known_lengths(ctx):
    """moira_filter_device on a fixed-pitch ragged slab (C3's layout): with max_length / min_length the first pass is sized
    for the decision (no ladder in decision mode); results are the oracle's either way."""
    slab, off, ln = synth.generate("v3v4", 50000, 51)
    stride = int(off[1] - off[0])
    dev = torch.device("cuda", 0)
    d_slab, d_len = torch.from_numpy(slab).to(dev), torch.from_numpy(ln.astype(np.int32)).to(dev)
    ee_o, ns_o = po.pb_batch(slab, off, ln, 0.005)
    ok_o = (ee_o + ns_o) <= ln * 0.01
    stream = torch.cuda.current_stream().cuda_stream
    for known in (True, False):
        for exact in (False, True):
            ee, ns, fl, cnt = _dev_arrays(len(ln))
            p = FilterParams(exact_ee=exact, max_length=int(ln.max()) if known else 0, min_length=int(ln.min()) if known else 0)
            ctx.filter_device(d_slab.data_ptr(), None, d_len.data_ptr(), stride, 0, len(ln), p, ee.data_ptr(), ns.data_ptr(),
                              fl.data_ptr(), cnt.data_ptr(), stream)
            torch.cuda.synchronize()
            f = fl.cpu().numpy()
            lb = (f & L.FLAG_LOWER_BOUND) != 0
            assert np.array_equal((f & 1) != 0, ok_o) and np.array_equal(ns.cpu().numpy(), ns_o)
            assert np.array_equal(ee.cpu().numpy()[~lb], ee_o[~lb]) and not (exact and lb.any())
            assert not (f & L.FLAG_NUMERIC).any()


def test_ccs_20k_reads(ctx):
    """C4-shaped reads (1 500 bp, Q up to 93, j* up to several hundred): 20 000 reads, ee / Ns bit-exact, decisions identical."""
    slab, off, ln = synth.generate("ccs", 20000, 61)
    ee_o, ns_o = po.pb_batch(slab, off, ln, 0.005)
    ok_o = (ee_o + ns_o) <= ln * 0.01
    res = ctx.filter_batch(slab, off, ln, FilterParams(exact_ee=True))
    assert np.array_equal(res.ee, ee_o) and np.array_equal(res.ns, ns_o) and np.array_equal(res.accept, ok_o)
    dec = ctx.filter_batch(slab, off, ln, FilterParams(exact_ee=False))
    lb = dec.lower_bound
    assert np.array_equal(dec.accept, ok_o) and np.array_equal(dec.ee[~lb], ee_o[~lb]) and (dec.ee[lb] <= ee_o[lb]).all()


def test_real_profile_generator_host_and_device(ctx):
    """The bootstrapped real-fixture workload: device generator -> filter == oracle on the same bytes."""
    dev = torch.device("cuda", 0)
    slab_t, lens, _ = synth.generate_device("real", 30000, 71, dev)
    assert lens is None
    slab = slab_t.cpu().numpy().reshape(-1)
    off = np.arange(30000, dtype=np.uint64) * np.uint64(256)
    ln = np.full(30000, 253, np.uint32)
    ee_o, ns_o = po.pb_batch(slab, off, ln, 0.005)
    res = ctx.filter_batch(slab, off, ln, FilterParams(exact_ee=True))
    assert np.array_equal(res.ee, ee_o) and np.array_equal(res.ns, ns_o)
    dec = ctx.filter_batch(slab, off, ln, FilterParams(exact_ee=False))
    assert np.array_equal(dec.accept, (ee_o + ns_o) <= 2.5300000000000002)
    assert 0.7 < dec.accept.mean() < 0.95


def _random_sequences(rng, n, pool, lo, hi):
    """n sequences drawn (with repeats) from `pool` distinct random strings of lo..hi bases, some differing only in the
    last base or in length."""
    base = []
    for _ in range(pool):
        L_ = int(rng.integers(lo, hi + 1))
        base.append(bytes(rng.choice(np.frombuffer(b"ACGTN", dtype=np.uint8), L_)))
    base += [b[:-1] + b"T" for b in base[:pool // 4]] + [b[:-1] for b in base[:pool // 4] if len(b) > 1]
    return [base[int(i)] for i in rng.integers(0, len(base), n)]


@pytest.mark.parametrize("lo,hi,truncate", [(1, 40, 0), (200, 260, 0), (1400, 1600, 0), (100, 300, 150)])
def test_device_dereplication_equals_host_collapse(ctx, lo, hi, truncate):
    """moira_collapse_device + moira_collapse_labels == moira_collapse (hash + memcmp on the host): groups, representatives,
    names order, abundance order -- on unaligned back-to-back sequences and on 16-byte rows."""
    rng = np.random.default_rng(lo + hi)
    n = 30000
    seqs = _random_sequences(rng, n, 3000, lo, hi)
    ee = rng.integers(0, 4, n).astype(np.float64) + rng.integers(0, 2, n) * 0.25
    ln = np.array([len(s_) for s_ in seqs], np.uint32)
    off = np.zeros(n, np.uint64)
    off[1:] = np.cumsum(ln[:-1], dtype=np.uint64)
    blob = np.frombuffer(b"".join(seqs) + b"\0" * 16, dtype=np.uint8)
    eff = np.minimum(ln, truncate) if truncate else ln
    want = moira_b200.collapse(blob, off, eff, ee)
    dev = torch.device("cuda", 0)
    d_blob, d_off, d_len = torch.from_numpy(blob.copy()).to(dev), torch.from_numpy(off.astype(np.int64)).to(dev), torch.from_numpy(ln.astype(np.int32)).to(dev)
    d_lab = torch.zeros(n, dtype=torch.int32, device=dev)
    ctx.collapse_device(d_blob.data_ptr(), d_off.data_ptr(), d_len.data_ptr(), 0, 0, n, d_lab.data_ptr(), truncate,
                        torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    labels = d_lab.cpu().numpy().view(np.uint32)
    got = moira_b200.collapse_labels(labels, ee)
    for f in ("group_of_read", "rep", "size", "member_start", "members", "order"):
        assert np.array_equal(getattr(got, f), getattr(want, f)), f
    # equal label <=> equal (truncated) sequence
    key = {}
    for r, s_ in enumerate(seqs):
        s_ = s_[:truncate] if truncate else s_
        assert key.setdefault(s_, int(labels[r])) == int(labels[r])
    assert len(set(key.values())) == len(key)


def test_fastq_ex_offsets_and_device_labels(ctx):
    """moira_filter_fastq_ex: record offsets == the host parser's, labels + ee -> the groups of the host collapse, with and
    without --truncate, on a text of several chunks."""
    import gzip, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = gzip.open(os.path.join(root, "tests", "golden", "test1.fastq.gz"), "rb").read() * 150      # ~ 90 MB: two chunks
    slab, off, ln, hoff, hlen, soff, qoff = moira_b200.parse_fastq(text, 33, True)
    for trunc in (None, 180):
        p = FilterParams(exact_ee=True, ee_output="final", truncate=trunc)
        r = ctx.filter_fastq_ex(text, p, 33, offsets=True, labels=True)
        assert np.array_equal(r.lengths, ln) and np.array_equal(r.seq_off, soff) and np.array_equal(r.qual_off, qoff)
        ref = ctx.filter_batch(slab, off, ln, p)
        assert np.array_equal(r.filter.ee, ref.ee) and np.array_equal(r.filter.flags, ref.flags)
        eff = np.minimum(ln, trunc) if trunc else ln
        want = moira_b200.collapse(text, soff, eff, ref.ee)
        got = moira_b200.collapse_labels(r.labels, r.filter.ee)
        for f in ("group_of_read", "rep", "size", "member_start", "members", "order"):
            assert np.array_equal(getattr(got, f), getattr(want, f)), f
        ho, hl = moira_b200.fastq_headers(text, r.seq_off)
        assert np.array_equal(ho, hoff) and np.array_equal(hl, hlen)


@pytest.mark.parametrize("extra", [[], ["-c", "False", "-o", "fastq", "-pi", "USEARCH", "-t", "200"], ["-l", "seq", "-me", "3"]])
def test_cli_shards_over_several_contexts_give_the_same_files(tmp_path, extra):
    """--devices: the reads of one FASTQ sharded over three contexts (own streams, parsers, dereplication tables) -- the
    shard labels are merged on the host -- write byte-identical files to the single-context run."""
    import gzip, os
    from moira_b200 import cli
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    fq = tmp_path / "in.fastq"
    fq.write_bytes(gzip.open(os.path.join(root, "tests", "golden", "test1.fastq.gz"), "rb").read() * 3)
    one, three = str(tmp_path / "one"), str(tmp_path / "three")
    assert cli.run(["-ffq", str(fq), "-op", one, "--silent", "--device", "0"] + extra) == 0
    assert cli.run(["-ffq", str(fq), "-op", three, "--silent", "--devices", "0,0,0"] + extra) == 0
    names = sorted(f[len("one"):] for f in os.listdir(tmp_path) if f.startswith("one."))
    assert len(names) >= 2
    for suffix in names:
        assert open(one + suffix, "rb").read() == open(three + suffix, "rb").read(), suffix


@pytest.mark.parametrize("profile,n", [("v3v4", 3_000_000), ("mixed", 3_000_000), ("real", 2_600_000)])
def test_multi_stage_cascade_equals_single_sweep(ctx, profile, n):
    """Decisions that need 5 .. 8 entries: the first stage picked on the device among 2 .. 5 entries (two pilots + verdicts)
    writes, array for array, what the single k-entry sweep writes; a sample is the oracle's."""
    dev = torch.device("cuda", 0)
    slab, lens, _ = synth.generate_device(profile, n, 81, dev)
    stride, fixed = synth.DEVICE_LAYOUT[profile]
    d_len = None if lens is None else lens.data_ptr()
    max_len = fixed or int(lens.max().item())
    min_len = fixed or int(lens.min().item())
    stream = torch.cuda.current_stream().cuda_stream
    marks = torch.zeros(n, dtype=torch.int32, device=dev)
    ctx.count_marks_device(slab.data_ptr(), None, d_len, stride, fixed or 0, n, marks.data_ptr(), 0, stream)
    # uncert 0.012 on 'real' (253 bp): cutoff 3.036 -> 5 entries, the smallest decision the multi-stage cascade takes
    uncert = 0.012 if profile == "real" else 0.01
    m = 3000
    h_slab = slab[:m].cpu().numpy().reshape(-1)
    off = np.arange(m, dtype=np.uint64) * np.uint64(stride)
    ln = np.full(m, fixed, np.uint32) if lens is None else lens[:m].cpu().numpy().astype(np.uint32)
    ee_o, ns_o = po.pb_batch(h_slab, off, ln, 0.005)
    for exact in (False, True):
        outs = []
        for cascade in (0, 2):
            ee, ns, fl, cnt = _dev_arrays(n)
            p = FilterParams(exact_ee=exact, cascade=cascade, max_length=max_len, min_length=min_len, length_sort=2, uncert=uncert)
            ctx.filter_device(slab.data_ptr(), None, d_len, stride, fixed or 0, n, p, ee.data_ptr(), ns.data_ptr(), fl.data_ptr(),
                              cnt.data_ptr(), stream, marks.data_ptr())
            torch.cuda.synchronize()
            outs.append((ee.cpu().numpy(), ns.cpu().numpy(), fl.cpu().numpy(), cnt.cpu().numpy()))
        for k in range(3):
            assert np.array_equal(outs[0][k], outs[1][k]), (exact, k)
        assert _same_but_diagnostics(outs[0][3], outs[1][3])
        if not exact:
            assert outs[0][3][L.CNT_FP64_OPS] < outs[1][3][L.CNT_FP64_OPS]          # the cascade executed less
        # oracle on a sample
        f = outs[0][2][:m]
        lb = (f & L.FLAG_LOWER_BOUND) != 0
        assert not (exact and lb.any()) and not (f & L.FLAG_NUMERIC).any()
        assert np.array_equal((f & 1) != 0, (ee_o + ns_o) <= ln * uncert)
        assert np.array_equal(outs[0][0][:m][~lb], ee_o[~lb]) and np.array_equal(outs[0][1][:m], ns_o)


@pytest.mark.parametrize("n,pool,ties", [(1, 1, False), (50_000, 7, True), (200_000, 60_000, False), (300_000, 300_000, True)])
def test_groups_on_the_device_equal_groups_on_the_host(ctx, n, pool, ties):
    """moira_collapse_labels_device == moira_collapse_labels: first-appearance numbering, representatives (first read with the
    strictly smallest ee), names order (record breakers to the front), abundance order with ties -- on a few huge groups,
    on mostly singletons, with many equal ee values."""
    rng = np.random.default_rng(n + pool)
    key = rng.integers(0, pool, n)
    _, first_idx, inv = np.unique(key, return_index=True, return_inverse=True)
    # the label is ANY member of the group (the read that claimed the hash slot), not necessarily the first
    pick = {}
    order = rng.permutation(n)
    for r in order[: min(n, 400_000)]:
        pick.setdefault(int(inv[r]), int(r))
    labels = np.array([pick[int(g)] for g in inv], dtype=np.uint32)
    ee = rng.integers(0, 5, n).astype(np.float64) if ties else rng.random(n) * 30
    want = moira_b200.collapse_labels(labels, ee)
    got = ctx.collapse_labels(labels, ee)
    for f in ("group_of_read", "rep", "size", "member_start", "members", "order"):
        assert np.array_equal(getattr(got, f), getattr(want, f)), f
    # the uint32 views, from host arrays and from device arrays
    d_lab, d_ee = torch.from_numpy(labels.view(np.int32)).to("cuda:0"), torch.from_numpy(ee).to("cuda:0")
    for views in (ctx.collapse_groups(labels, ee), ctx.collapse_groups(d_lab.data_ptr(), d_ee.data_ptr(), n)):
        for f in ("group_of_read", "rep", "size", "member_start", "members", "order"):
            assert getattr(views, f).dtype == np.uint32 and np.array_equal(getattr(views, f), getattr(want, f)), f
    with pytest.raises(moira_b200.MoiraError):
        bad = labels.copy()
        bad[n // 2] = n + 5
        ctx.collapse_labels(bad, ee)


@pytest.mark.parametrize("lo,hi,truncate", [(1, 60, 0), (240, 260, 0), (100, 300, 120)])
def test_labels_of_host_sequences_by_address(ctx, lo, hi, truncate):
    """moira_collapse_addr: sequences scattered over several host buffers (as the contig rows of several batches are) ->
    labels on the device -> groups == the host's hash + memcmp collapse of the same strings."""
    rng = np.random.default_rng(hi * 7 + truncate)
    n = 40000
    seqs = _random_sequences(rng, n, 4000, lo, hi)
    ee = rng.integers(0, 6, n).astype(np.float64) * 0.5
    ln = np.array([len(s_) for s_ in seqs], np.uint32)
    # three separate buffers with gaps
    bufs, addr = [], np.zeros(n, np.uint64)
    for part in range(3):
        idx = np.arange(part, n, 3)
        blob = np.frombuffer(b"".join(seqs[i] + b"##" for i in idx), dtype=np.uint8).copy()
        bufs.append(blob)
        pos = np.zeros(len(idx), np.uint64)
        pos[1:] = np.cumsum(ln[idx][:-1].astype(np.uint64) + np.uint64(2))
        addr[idx] = pos + np.uint64(blob.ctypes.data)
    labels = ctx.collapse_addr(addr, ln, truncate)
    eff = np.minimum(ln, truncate) if truncate else ln
    want = moira_b200.collapse(None, addr, eff, ee)
    got = ctx.collapse_groups(labels, ee)
    for f in ("group_of_read", "rep", "size", "member_start", "members", "order"):
        assert np.array_equal(getattr(got, f), getattr(want, f)), f


def _fresh_ctx(env):
    """A context created under environment switches (the library reads them at context creation)."""
    import os
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        return moira_b200.Context(0)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("n,maxerrors,length_sort", [(60000, None, 2), (60000, None, 1), (3000, None, 0), (40000, 24.0, 0)])
def test_classify_first_equals_first_pass_with_the_decisions_k(n, maxerrors, length_sort):
    """Long reads (1 500 bp: the decision needs 17 entries): by default the fp32 classifier runs first and the ladder sweeps
    every read once with its own K.  Against the route with a K = 18 first pass: exact mode writes the same arrays and
    counters; decision mode the same decisions, the same ee for every accepted read, valid lower bounds for the others;
    both are the oracle's.  length_sort = 1: the classifier walks the length-sorted permutation."""
    slab, off, ln = synth.generate("ccs", n, 91)
    slab = slab.copy()
    rng = np.random.default_rng(5)
    slab[off[rng.integers(0, n, 500)] + rng.integers(0, 1500, 500).astype(np.uint64)] = 0xFF      # some N
    ln = ln.copy()
    ln[:50] = rng.integers(0, 40, 50)                                                                # and some very short / empty reads
    ee_o, ns_o = po.pb_batch(slab, off, ln, 0.005)
    cut = np.full(n, maxerrors) if maxerrors else ln * 0.01
    ok_o = (ee_o + ns_o) <= cut
    kw = dict(maxerrors=maxerrors, length_sort=length_sort) if maxerrors else dict(length_sort=length_sort)
    cf = _fresh_ctx({})
    fp = _fresh_ctx({"MOIRA_B200_CLASSIFY_FIRST_K": "1000", "MOIRA_B200_CLASSIFY_FIRST_DEC_K": "1000"})
    try:
        for exact in (True, False):
            p = FilterParams(exact_ee=exact, **kw)
            a, b = cf.filter_batch(slab, off, ln, p), fp.filter_batch(slab, off, ln, p)
            assert int(a.counters[L.CNT_CLASSIFIED]) == n and int(b.counters[L.CNT_CLASSIFIED]) == 0
            assert np.array_equal(a.accept, ok_o) and np.array_equal(b.accept, ok_o)
            assert np.array_equal(a.ns, ns_o) and np.array_equal(b.ns, ns_o)
            assert not (a.flags & L.FLAG_NUMERIC).any()
            if exact:
                assert np.array_equal(a.ee, ee_o) and np.array_equal(b.ee, ee_o) and np.array_equal(a.flags, b.flags)
                assert _same_but_diagnostics(a.counters, b.counters)
            else:
                la, lb = a.lower_bound, b.lower_bound
                assert np.array_equal(a.ee[~la], ee_o[~la]) and (a.ee[la] <= ee_o[la]).all() and not (la & a.accept).any()
                assert np.array_equal(b.ee[~lb], ee_o[~lb]) and (b.ee[lb] <= ee_o[lb]).all()
                keep = [L.CNT_READS, L.CNT_ACCEPTED, L.CNT_BAD_ERRORS, L.CNT_BAD_LENGTH, L.CNT_BAD_AMBIGS]
                assert np.array_equal(np.asarray(a.counters)[keep], np.asarray(b.counters)[keep])
                # fewer FP64 operations than sweeping 18 entries for every read
                assert int(a.counters[L.CNT_FP64_OPS]) < int(b.counters[L.CNT_FP64_OPS])
    finally:
        cf.close()
        fp.close()


def test_length_bucketed_exact_first_pass_with_capped_k():
    """Ragged batches in exact mode: the length-bucketed first pass sweeps at most four entries and the ladder does the
    rest; same arrays and counters as with every bucket's own decision K, and the oracle's."""
    n = 80000
    slab, off, ln = synth.generate("mixed", n, 93)
    ee_o, ns_o = po.pb_batch(slab, off, ln, 0.005)
    capped = _fresh_ctx({})
    full = _fresh_ctx({"MOIRA_B200_EXACT_SORTED_KCAP": "0"})
    try:
        p = FilterParams(exact_ee=True, length_sort=1)
        a, b = capped.filter_batch(slab, off, ln, p), full.filter_batch(slab, off, ln, p)
        assert np.array_equal(a.ee, ee_o) and np.array_equal(a.ns, ns_o) and np.array_equal(b.ee, ee_o)
        assert np.array_equal(a.flags, b.flags) and _same_but_diagnostics(a.counters, b.counters)
        assert int(a.counters[L.CNT_ESCALATED]) > int(b.counters[L.CNT_ESCALATED])
    finally:
        capped.close()
        full.close()


def test_length_bucketed_decision_with_capped_buckets(ctx):
    """Ragged batches whose decisions need 5 .. 8 entries: buckets swept with at most four entries first (Newton bound at every
    read's own K, the rest to the rung that holds it) -- by pilot (cascade = 0) or forced (1) -- give the decisions, Ns and
    accepted-read ee of the sweep with every bucket's own K (2); lower bounds are valid; a sample is the oracle's."""
    n = 3_000_000
    dev = torch.device("cuda", 0)
    slab, lens, _ = synth.generate_device("mixed", n, 97, dev)
    stride, fixed = synth.DEVICE_LAYOUT["mixed"]
    max_len, min_len = int(lens.max().item()), int(lens.min().item())
    stream = torch.cuda.current_stream().cuda_stream
    m = 4000
    idx = np.sort(np.random.default_rng(9).choice(n, m, replace=False))
    h_rows = slab[torch.as_tensor(idx, device=dev)].cpu().numpy().reshape(-1)
    off = np.arange(m, dtype=np.uint64) * np.uint64(stride)
    ln = lens[torch.as_tensor(idx, device=dev)].cpu().numpy().astype(np.uint32)
    ee_o, ns_o = po.pb_batch(h_rows, off, ln, 0.005)
    for ambigs in ("treat_as_errors", "ignore"):
        outs = []
        for cascade in (0, 1, 2):
            ee, ns, fl, cnt = _dev_arrays(n)
            p = FilterParams(exact_ee=False, cascade=cascade, max_length=max_len, min_length=min_len, length_sort=1, ambigs=ambigs)
            ctx.filter_device(slab.data_ptr(), None, lens.data_ptr(), stride, 0, n, p, ee.data_ptr(), ns.data_ptr(), fl.data_ptr(),
                              cnt.data_ptr(), stream)
            torch.cuda.synchronize()
            outs.append((ee.cpu().numpy(), ns.cpu().numpy(), fl.cpu().numpy(), cnt.cpu().numpy()))
        ref = outs[2]
        lb_ref = (ref[2] & L.FLAG_LOWER_BOUND) != 0
        for o in outs[:2]:
            lb = (o[2] & L.FLAG_LOWER_BOUND) != 0
            assert np.array_equal(o[2] & 0x0F, ref[2] & 0x0F)                    # accept bit and reason
            assert np.array_equal(o[1], ref[1])
            assert not ((o[2] & 1) & lb).any() and not (o[2] & L.FLAG_NUMERIC).any()
            both = ~lb & ~lb_ref
            assert np.array_equal(o[0][both], ref[0][both])
            keep = [L.CNT_READS, L.CNT_ACCEPTED, L.CNT_BAD_ERRORS, L.CNT_BAD_LENGTH, L.CNT_BAD_AMBIGS]
            assert np.array_equal(o[3][keep], ref[3][keep])
            assert int(o[3][L.CNT_ESCALATED]) > 0 and int(o[3][L.CNT_FP64_OPS]) < int(ref[3][L.CNT_FP64_OPS])
            # oracle on the sample
            f, e = o[2][idx], o[0][idx]
            s_lb = (f & L.FLAG_LOWER_BOUND) != 0
            cut = ln * 0.01
            stat = ee_o + (ns_o if ambigs == "treat_as_errors" else 0)
            assert np.array_equal((f & 1) != 0, stat <= cut)
            assert np.array_equal(e[~s_lb], ee_o[~s_lb]) and (e[s_lb] <= ee_o[s_lb]).all()       # ee_output = raw: the statistic without Ns


@pytest.mark.parametrize("kw", [dict(alpha=1e-6), dict(alpha=0.2), dict(uncert=0.02), dict(maxerrors=9.5), dict(ambigs="ignore"),
                                dict(ambigs="disallow"), dict(round=True), dict(truncate=1200), dict(truncate=900, maxerrors=12.0),
                                dict(uncert=0.05, alpha=0.05)])
def test_classify_first_decisions_under_every_rule(kw):
    """Classify-first in decision mode (Chernoff rejects included) under the decision rules of write_results: accept bit and
    reason equal those of a context that sweeps the decision's K first, accepted reads carry the same ee, lower bounds are
    below the exact statistic (taken from an exact-mode run of that context)."""
    n = 40000
    slab, off, ln = synth.generate("ccs", n, 101)
    slab = slab.copy()
    ln = ln.copy()
    rng = np.random.default_rng(11)
    slab[off[rng.integers(0, n, 3000)] + rng.integers(0, 1400, 3000).astype(np.uint64)] = 0xFF       # N in one read of eight
    slab[off[rng.integers(0, n, 300)] + rng.integers(0, 1400, 300).astype(np.uint64)] = 0xFE         # and some n
    ln[rng.integers(0, n, 2000)] = rng.integers(1, 1500, 2000)                                        # ragged
    cf = _fresh_ctx({})
    fp = _fresh_ctx({"MOIRA_B200_CLASSIFY_FIRST_K": "1000", "MOIRA_B200_CLASSIFY_FIRST_DEC_K": "1000"})
    try:
        for length_sort in (2, 1):
            exact = fp.filter_batch(slab, off, ln, FilterParams(exact_ee=True, length_sort=length_sort, **kw))
            a = cf.filter_batch(slab, off, ln, FilterParams(exact_ee=False, length_sort=length_sort, **kw))
            b = fp.filter_batch(slab, off, ln, FilterParams(exact_ee=False, length_sort=length_sort, **kw))
            assert int(a.counters[L.CNT_CLASSIFIED]) > 0 and int(b.counters[L.CNT_CLASSIFIED]) == 0
            assert np.array_equal(a.flags & 0x0F, b.flags & 0x0F) and np.array_equal(a.flags & 0x0F, exact.flags & 0x0F)
            assert np.array_equal(a.ns, exact.ns) and not (a.flags & L.FLAG_NUMERIC).any()
            la = a.lower_bound
            assert not (la & a.accept).any()
            assert np.array_equal(a.ee[~la], exact.ee[~la]) and (a.ee[la] <= exact.ee[la]).all()
            keep = [L.CNT_READS, L.CNT_ACCEPTED, L.CNT_BAD_ERRORS, L.CNT_BAD_LENGTH, L.CNT_BAD_AMBIGS]
            assert np.array_equal(np.asarray(a.counters)[keep], np.asarray(exact.counters)[keep])
    finally:
        cf.close()
        fp.close()
