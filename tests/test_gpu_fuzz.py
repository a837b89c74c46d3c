"""GPU (-m gpu): randomised batches x randomised decision rules through every route the library picks on its own (single sweep,
cascades, classify-first, length-bucketed passes, ladder, 6-bit transport) against the oracle.  Fixed seeds: failures reproduce."""
import numpy as np
import pytest

import moira_b200
from moira_b200 import FilterParams
from moira_b200 import _lib as L
from oracle import py_oracle as po

pytestmark = pytest.mark.gpu


def _batch(rng, n, shape):
    """Random reads in the in-band slab layout: (slab, offsets, lengths)."""
    if shape == "short":
        ln = rng.integers(0, 120, n)
    elif shape == "amplicon":
        ln = np.full(n, int(rng.integers(150, 500)))
    elif shape == "ragged":
        ln = rng.integers(100, 700, n)
    elif shape == "long":
        ln = rng.integers(900, 2200, n)
    else:   # "wild": everything at once, a few very long reads
        ln = np.where(rng.random(n) < 0.02, rng.integers(2000, 5000, n), rng.integers(0, 600, n))
    ln = ln.astype(np.uint32)
    stride = int((int(ln.max()) + 15) // 16 * 16) if n else 16
    stride = max(stride, 16)
    slab = np.full((n, stride), 0xFD, np.uint8)
    kind = rng.random(n)
    for r in range(n):
        l = int(ln[r])
        if not l:
            continue
        if kind[r] < 0.6:
            q = rng.integers(28, 42, l)
        elif kind[r] < 0.8:
            q = rng.integers(2, 25, l)
        elif kind[r] < 0.9:
            q = np.where(rng.random(l) < 0.15, rng.integers(0, 12, l), rng.integers(30, 94, l))
        else:
            q = rng.integers(0, 60, l)
        q = q.astype(np.uint8)
        m = rng.random(l)
        q[m < 0.004] = 0xFF
        q[(m >= 0.004) & (m < 0.005)] = 0xFE
        slab[r, :l] = q
    off = np.arange(n, dtype=np.uint64) * np.uint64(stride)
    return slab.reshape(-1), off, ln


CASES = [(seed, shape, n) for seed, (shape, n) in enumerate(
    [("short", 3000), ("amplicon", 3000), ("ragged", 3000), ("long", 1500), ("wild", 3000), ("ragged", 40000), ("short", 40000),
     ("amplicon", 6000), ("long", 2500), ("wild", 40000), ("amplicon", 40000), ("ragged", 5000)])]


@pytest.mark.parametrize("seed,shape,n", CASES)
def test_random_batches_and_rules_against_the_oracle(ctx, seed, shape, n):
    rng = np.random.default_rng(1000 + seed)
    slab, off, ln = _batch(rng, n, shape)
    for trial in range(4):
        alpha = float(rng.choice([1e-6, 1e-3, 0.005, 0.05, 0.2]))
        kw = dict(alpha=alpha, ambigs=str(rng.choice(["treat_as_errors", "ignore", "disallow"])), round=bool(rng.random() < 0.25),
                  length_sort=int(rng.choice([0, 1, 2])), cascade=int(rng.choice([0, 0, 1, 2])))
        if rng.random() < 0.35:
            kw["maxerrors"] = float(rng.choice([0.5, 2.0, 7.3, 15.0, 40.0]))
        else:
            kw["uncert"] = float(rng.choice([0.002, 0.01, 0.02, 0.05]))
        if rng.random() < 0.3:
            kw["truncate"] = int(rng.integers(1, max(2, int(ln.max()))))
        if rng.random() < 0.2:
            kw["slab_format"] = "q6"
        trunc = kw.get("truncate")
        eff = np.minimum(ln, trunc).astype(np.uint32) if trunc else ln
        ee_o, ns_o = po.pb_batch(slab, off, eff, alpha)
        has_n = np.array([(slab[int(o):int(o) + int(l)] == 0xFF).any() for o, l in zip(off, eff)], dtype=bool)
        ok_o, reason_o, _ = po.decide_batch(ee_o, ns_o, ln, has_n, thr_kind="maxerrors" if "maxerrors" in kw else "uncert",
                                            thr=kw.get("maxerrors", kw.get("uncert")), ambigs=kw["ambigs"], round_flag=kw["round"],
                                            truncate=trunc)
        use_q6 = kw.pop("slab_format", None) == "q6" and int(slab[slab < 0xFD].max(initial=0)) <= 60
        for exact in (True, False):
            p = FilterParams(exact_ee=exact, **kw)
            if use_q6:
                res = ctx.filter_batch(moira_b200.pack_q6(slab), off, ln, FilterParams(exact_ee=exact, slab_format="q6", **kw))
            else:
                res = ctx.filter_batch(slab, off, ln, p)
            tag = (seed, shape, trial, exact, kw)
            assert not (res.flags & L.FLAG_NUMERIC).any(), tag
            near = res.near_cutoff
            assert np.array_equal(res.accept[~near], ok_o[~near]) and np.array_equal(res.reason[~near], reason_o[~near]), tag
            assert np.array_equal(res.ns, ns_o), tag
            lb = res.lower_bound
            assert not (exact and lb.any()) and not (lb & res.accept).any(), tag
            assert np.array_equal(res.ee[~lb], ee_o[~lb]) and (res.ee[lb] <= ee_o[lb]).all(), tag
            assert int(res.counters[L.CNT_READS]) == n and int(res.counters[L.CNT_ACCEPTED]) == int(res.accept.sum()), tag


def test_single_read_calls_of_any_length_against_the_oracle():
    """The drop-in single-read functions (batches of one: every route, classify-first and the wide rungs included) on reads of
    1 .. 6 000 bases, clean and noisy, against the C restatement."""
    import bernoulli
    import moira
    rng = np.random.default_rng(77)
    for length in (1, 2, 17, 100, 301, 899, 1500, 3000, 6000):
        for lo, hi in ((30, 42), (2, 20), (0, 94)):
            quals = [int(q) for q in rng.integers(lo, hi, length)]
            seq = "".join(rng.choice(list("ACGTNn"), size=length, p=[.245, .245, .245, .245, .015, .005]))
            for alpha in (0.005, 0.1):
                want_c = po.pb_c(seq, [q if q > 0 else 1 for q in quals], alpha)
                assert bernoulli.calculate_errors_PB(seq, quals, alpha) == want_c, (length, lo, alpha)
                up = seq.replace("n", "A")                                  # the Python twin counts 'N' only: same answer without 'n'
                assert moira.calculate_errors_PB(up, quals, alpha) == po.pb_c(up, [q if q > 0 else 1 for q in quals], alpha)


@pytest.mark.parametrize("shape,n", [("long", 4000), ("ragged", 20000), ("wild", 6000)])
def test_fastq_text_through_the_device_parser_equals_the_slab_path(ctx, shape, n):
    """moira_filter_fastq (text -> device parser -> whatever route the filter picks: classify-first for the long reads, the
    bucketed passes for the ragged ones) == moira_filter_batch on the slab the host packer makes of the same reads, both modes."""
    rng = np.random.default_rng(500 + n)
    slab, off, ln = _batch(rng, n, shape)
    ln = np.maximum(ln, 1).astype(np.uint32)                      # FASTQ records cannot be empty (EmptySeqError)
    recs = []
    for r in range(n):
        q = slab[int(off[r]):int(off[r]) + int(ln[r])]
        seq = np.where(q == 0xFF, ord("N"), np.where(q == 0xFE, ord("n"), ord("A"))).astype(np.uint8).tobytes()
        qual = (np.where(q >= 0xFD, 2, np.minimum(q, 93)) + 33).astype(np.uint8).tobytes()
        recs.append(b"@r%d\n%s\n+\n%s\n" % (r, seq, qual))
    text = b"".join(recs)
    pslab, poff, pln, *_ = moira_b200.parse_fastq(text, 33, True)
    for exact in (True, False):
        p = FilterParams(exact_ee=exact)
        want = ctx.filter_batch(pslab, poff, pln, p)
        got, glen = ctx.filter_fastq(text, p)
        assert np.array_equal(glen, pln)
        assert np.array_equal(got.ns, want.ns) and np.array_equal(got.flags & 0x0F, want.flags & 0x0F)
        lb = got.lower_bound | want.lower_bound
        assert np.array_equal(got.ee[~lb], want.ee[~lb]) and not (exact and lb.any())
        assert not (got.flags & L.FLAG_NUMERIC).any()
