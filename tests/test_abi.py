"""CPU: the C-ABI library loads, exports what include/moira_b200.h declares, its host-side pieces
(tables, packers, parsers, validation) agree with the oracle, and it fails loudly without a GPU."""
import ctypes
import gzip
import os
import re

import numpy as np
import pytest

import moira_b200
from moira_b200 import _lib as L
from oracle import py_oracle as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_all_exported():
    hdr = open(os.path.join(ROOT, "include", "moira_b200.h")).read()
    declared = re.findall(r"^MOIRA_API [^;(]*?\b(moira_\w+)\(", hdr, re.M)
    assert len(declared) >= 20
    assert sorted(declared) == sorted(L.EXPORTS)
    for name in declared:
        assert hasattr(L.lib, name), name
    assert L.lib.moira_abi_version() == int(re.search(r"#define MOIRA_ABI_VERSION (\d+)", hdr).group(1))


def test_header_constants_match_binding():
    hdr = open(os.path.join(ROOT, "include", "moira_b200.h")).read()
    consts = dict(re.findall(r"^#define MOIRA_(\w+)\s+(-?(?:0x)?[0-9A-Fa-f]+)\b", hdr, re.M))
    for k, v in consts.items():
        if hasattr(L, k):
            assert getattr(L, k) == int(v, 0), k
    assert ctypes.sizeof(L.Params) == 64
    # the ctypes mirror of moira_params has the header's fields, in the header's order
    body = re.search(r"typedef struct moira_params \{(.*?)\} moira_params;", hdr, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"\b(?:int32_t|uint32_t|double)\s+(\w+)\s*;", body)
    assert fields == [f[0] for f in L.Params._fields_] and "cascade" in fields


def test_lut_matches_oracle_tables():
    p, q, e, eqp = moira_b200.build_lut()
    op, oq, oe = np.zeros(256), np.zeros(256), np.zeros(256)
    po.oracle_lib().oracle_tables(op.ctypes.data, oq.ctypes.data, oe.ctypes.data)
    assert np.array_equal(p[1:0xFD], op[1:0xFD])
    assert np.array_equal(q[1:0xFD], oq[1:0xFD])
    assert np.array_equal(e[1:0xFD], oe[1:0xFD])
    assert (p[0], q[0], e[0]) == (op[1], oq[1], oe[1])             # Q == 0 -> 1
    assert np.all(p[0xFD:] == 0) and np.all(q[0xFD:] == 1) and np.all(e[0xFD:] == 0)
    assert eqp is True                                             # glibc: e == p bitwise


def test_pack_reads_matches_oracle_packer(forward_records):
    seqs = [s for _, s, _ in forward_records[:200]] + ["ACGTNnacgt", "N", "a"]
    quals = [q for _, _, q in forward_records[:200]] + [[0, 1, 2, 3, 4, 5, 93, 252, 40, -3], [7], [9]]
    for lower in (True, False):
        slab, off, ln = moira_b200.pack_reads(seqs, quals, lower_n_ambiguous=lower)
        qs = [[max(v, 0) for v in q] for q in quals]
        oslab, ooff, oln = po.pack_records(seqs, qs, lower_n_ambiguous=lower)
        assert np.array_equal(off, ooff) and np.array_equal(ln, oln)
        assert np.array_equal(slab, oslab[:slab.size])


def test_pack_rejects_unrepresentable_quality():
    with pytest.raises(moira_b200.MoiraError) as ei:
        moira_b200.pack_reads(["AC"], [[10, 253]])
    assert ei.value.code == L.ERR_BAD_QUALITY


def test_parse_fastq_matches_oracle_parser(forward_records):
    text = gzip.open(os.path.join(ROOT, "tests", "golden", "test1.fastq.gz"), "rb").read()
    slab, off, ln, hoff, hlen, soff, qoff = moira_b200.parse_fastq(text, 33, True)
    assert len(ln) == len(forward_records) == 1000
    seqs = [s for _, s, _ in forward_records]
    quals = [q for _, _, q in forward_records]
    oslab, ooff, oln = po.pack_records(seqs, [[max(v, 0) for v in q] for q in quals])
    assert np.array_equal(off, ooff) and np.array_equal(ln, oln) and np.array_equal(slab, oslab[:slab.size])
    for i in (0, 1, 17, 999):
        h = text[int(hoff[i]):int(hoff[i]) + int(hlen[i])].decode().replace(":", "_")
        assert h == forward_records[i][0]
        assert text[int(soff[i]):int(soff[i]) + int(ln[i])].decode() == seqs[i]


def test_parse_fastq_errors():
    for bad, word in ((b"@r1\nACGT\n+\nIII\n", "LengthMismatchError"), (b"@r1\n\n+\nIII\n", "EmptySeqError"),
                      (b"@r1\nACG\n+\n\n", "EmptyQualError")):
        with pytest.raises(moira_b200.MoiraError) as ei:
            moira_b200.parse_fastq(bad)
        assert ei.value.code == L.ERR_PARSE and word in ei.value.message
    slab, off, ln, *_ = moira_b200.parse_fastq(b"@r1 x\tb\n ACGN \n+\nII#I\n@partial\nAC\n")
    assert list(ln) == [4] and list(slab[:4]) == [40, 40, 2, 0xFF]


def test_shim_validation_is_the_reference_binding_s():
    # bernoullimodule.c:74-90: TypeError for wrong types, ValueError for alpha / length -- raised
    # before any GPU work, so checkable on CPU
    from moira_b200 import bernoulli
    with pytest.raises(ValueError, match="Alpha must be between 0 and 1"):
        bernoulli.calculate_errors_PB("ACGT", [30] * 4, 0.0)
    with pytest.raises(ValueError, match="Alpha must be between 0 and 1"):
        bernoulli.calculate_errors_PB("ACGT", [30] * 4, 1.0)
    with pytest.raises(ValueError, match="same length"):
        bernoulli.calculate_errors_PB("ACGT", [30] * 3, 0.005)
    with pytest.raises(TypeError):
        bernoulli.calculate_errors_PB("ACGT", (30, 30, 30, 30), 0.005)
    with pytest.raises(TypeError):
        bernoulli.calculate_errors_PB(b"ACGT", [30] * 4, 0.005)
    with pytest.raises(TypeError):
        bernoulli.calculate_errors_PB("ACGT", [30, 30, 3.5, 30], 0.005)
    import bernoulli as top
    assert top.calculate_errors_PB is bernoulli.calculate_errors_PB


def test_no_silent_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(moira_b200.MoiraError) as ei:
        moira_b200.Context(0)
    assert ei.value.code == L.ERR_CUDA


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "moira_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                for needle in ("import oracle", "from oracle", "liboracle", "oracle/", "py_oracle", "_ref"):
                    assert needle not in src, (f, needle)


def test_parse_fastq_multithreaded_equals_single_thread():
    """>= 1 MB of FASTQ takes the parallel path (ranges cut anywhere, also inside records whose
    quality line starts with '@'); results must equal the one-thread parse."""
    rng = np.random.default_rng(12)
    recs = []
    for i in range(6000):
        L_ = int(rng.integers(1, 400))
        seq = "".join(rng.choice(list("ACGTNn"), size=L_, p=[.24, .24, .24, .24, .03, .01]))
        q = rng.integers(33, 74, size=L_)
        if i % 3 == 0:
            q[0] = ord("@")                       # '@' as a quality character must not start a record
        if i % 5 == 0:
            q[-1] = ord("+")
        recs.append("@r%d:x desc\n%s\n+r%d\n%s\n" % (i, seq, i, q.astype(np.uint8).tobytes().decode()))
    text = "".join(recs).encode() + b"@partial\nACGT\n+\n"
    assert len(text) > (1 << 20)
    outs = []
    for threads in (1, 7, 16):
        assert L.lib.moira_set_host_threads(threads) == 0
        outs.append(moira_b200.parse_fastq(text, 33, True))
    L.lib.moira_set_host_threads(0)
    assert len(outs[0][2]) == 6000
    rows0 = [outs[0][0][int(o):int(o) + int(l)] for o, l in zip(outs[0][1], outs[0][2])]
    for other in outs[1:]:
        for a, b in zip(outs[0][2:], other[2:]):            # lengths and the byte ranges inside the text
            assert np.array_equal(a, b)
        assert np.all(other[1] % 16 == 0) and np.all(np.diff(other[1].astype(np.int64)) > 0)
        for r0, o, l in zip(rows0, other[1], other[2]):     # same rows, possibly with slack between threads' slices
            assert np.array_equal(r0, other[0][int(o):int(o) + int(l)])
    # errors are reported for the first bad record, whatever the thread count
    bad = "".join(recs[:3000]).encode() + b"@bad\nACGT\n+\nIII\n" + "".join(recs[3000:]).encode()
    for threads in (1, 9):
        L.lib.moira_set_host_threads(threads)
        with pytest.raises(moira_b200.MoiraError) as ei:
            moira_b200.parse_fastq(bad, 33, True)
        assert ei.value.code == L.ERR_PARSE and "record 3000" in ei.value.message and "LengthMismatchError" in ei.value.message
    L.lib.moira_set_host_threads(0)


def test_collapse_matches_reference_semantics(forward_records, forward_names):
    """moira_collapse vs the oracle's restatement of moira.py:459-475/491-504: groups, representatives,
    names order and abundance order -- on the forward fixture (golden .names) and on synthetic duplicates."""
    args = po.Args()
    ee = np.array([po.process_filter(s, q, args)[2] for _, s, q in forward_records])
    seqs = [s for _, s, _ in forward_records]
    heads = [h for h, _, _ in forward_records]
    ln = np.array([len(s) for s in seqs], dtype=np.uint32)
    off = np.zeros(len(seqs), dtype=np.uint64)
    off[1:] = np.cumsum(ln[:-1])
    for threads in (1, 5):
        col = moira_b200.collapse("".join(seqs).encode(), off, ln, ee, n_threads=threads)
        got = {}
        for g in col.order.tolist():
            m = col.members[int(col.member_start[g]):int(col.member_start[g + 1])]
            got[heads[int(col.rep[g])]] = [heads[int(r)] for r in m]
            assert int(col.size[g]) == len(m) and int(m[0]) == int(col.rep[g])
        golden = dict(forward_names["good"])
        golden.update(forward_names["bad"])
        assert got == golden and len(got) == 487
        assert np.all(np.diff(col.size[col.order].astype(np.int64)) <= 0)        # abundance, largest first
    # synthetic: heavy duplication, equal-ee ties, one-read groups; compare with the pure-python loop
    rng = np.random.default_rng(4)
    pool = ["".join(rng.choice(list("ACGT"), size=int(rng.integers(5, 40)))) for _ in range(300)]
    seqs = [pool[int(i)] for i in rng.zipf(1.3, size=20000) % 300]
    ee = np.round(rng.random(20000) * 4, 1)                                       # many exact ties
    uniq = {}
    for r, (s, e) in enumerate(zip(seqs, ee)):
        u = uniq.get(s)
        if u is None:
            uniq[s] = [r, e, [r]]
        elif e < u[1]:
            u[0], u[1] = r, e
            u[2].insert(0, r)
        else:
            u[2].append(r)
    order = sorted(uniq, key=lambda s: len(uniq[s][2]), reverse=True)
    ln = np.array([len(s) for s in seqs], dtype=np.uint32)
    off = np.zeros(len(seqs), dtype=np.uint64)
    off[1:] = np.cumsum(ln[:-1])
    col = moira_b200.collapse("".join(seqs).encode(), off, ln, ee)
    assert len(col.order) == len(order)
    for g, s in zip(col.order.tolist(), order):
        assert int(col.rep[g]) == uniq[s][0]
        assert col.members[int(col.member_start[g]):int(col.member_start[g + 1])].tolist() == uniq[s][2]
    assert np.array_equal(col.group_of_read[col.rep.astype(np.int64)], np.arange(len(col.rep), dtype=np.uint64))


def test_q6_transport_image_roundtrip():
    """moira_pack_q6: 16 slab bytes -> 12 image bytes; decoded with numpy it gives the slab back."""
    rng = np.random.default_rng(1)
    slab = rng.integers(0, 61, size=16 * 70000, dtype=np.uint8)
    slab[rng.random(slab.size) < 0.02] = 0xFF
    slab[rng.random(slab.size) < 0.01] = 0xFE
    slab[rng.random(slab.size) < 0.03] = 0xFD
    for threads in (1, 6):
        img = moira_b200.pack_q6(slab, n_threads=threads)
        assert img.size == slab.size * 3 // 4
        b = img.reshape(-1, 3).astype(np.uint32)
        v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
        codes = np.stack([(v >> (6 * k)) & 63 for k in range(4)], axis=1).reshape(-1)
        back = np.where(codes > 60, codes + 192, codes).astype(np.uint8)
        assert np.array_equal(back, slab)
    bad = slab.copy()
    bad[12345] = 61
    with pytest.raises(moira_b200.MoiraError) as ei:
        moira_b200.pack_q6(bad)
    assert ei.value.code == L.ERR_BAD_QUALITY


def _fasta_qual_text(records):
    fa = "".join(">%s some description\n%s\n" % (h.replace("_", ":"), s) for h, s, _ in records)
    qu = "".join(">%s\tother words\n%s\n" % (h.replace("_", ":"), " ".join(map(str, q))) for h, _, q in records)
    return fa.encode(), qu.encode()


def test_parse_fasta_qual_matches_oracle_parser(forward_records):
    """moira_parse_fasta_qual against the oracle's restatement of moira.py:1093-1149 on the reference's own
    test reads written as fasta + qual."""
    fa, qu = _fasta_qual_text(forward_records)
    recs = po.parse_fasta_qual_text(fa.decode(), qu.decode())
    assert [(h, s, q) for h, s, q in recs] == [(h, s, q) for h, s, q in forward_records]
    slab, qslab, off, ln, hoff, hlen, soff = moira_b200.parse_fasta_qual(fa, qu, True)
    oslab, ooff, oln = po.pack_records([r[1] for r in recs], [[max(v, 0) for v in r[2]] for r in recs])
    assert np.array_equal(off, ooff) and np.array_equal(ln, oln) and np.array_equal(slab, oslab[:slab.size])
    for i in (0, 3, 500, 999):
        assert fa[int(hoff[i]):int(hoff[i]) + int(hlen[i])].decode().replace(":", "_") == recs[i][0]
        assert fa[int(soff[i]):int(soff[i]) + int(ln[i])].decode() == recs[i][1]
        assert list(qslab[int(off[i]):int(off[i]) + int(ln[i])]) == [max(v, 0) for v in recs[i][2]]


def test_parse_fasta_qual_errors_and_threads():
    P = moira_b200.parse_fasta_qual
    for fa, qu, word in ((b">a\nACGT\n", b">b\n1 2 3 4\n", "NameMismatchError"), (b">a\n\n", b">a\n1 2\n", "EmptySeqError"),
                         (b">a\nAC\n", b">a\n\n", "EmptyQualError"), (b">a\nACG\n", b">a\n1 2\n", "LengthMismatchError"),
                         (b">a\nACG\n", b">a\n1 x 3\n", "ValueError"), (b">a\nAC\n>b\nAC\n", b">a\n1 2\n", "NameMismatchError")):
        with pytest.raises(moira_b200.MoiraError) as ei:
            P(fa, qu)
        assert ei.value.code == L.ERR_PARSE and word in ei.value.message, (fa, qu)
    with pytest.raises(moira_b200.MoiraError) as ei:
        P(b">a\nAC\n", b">a\n7 300\n")
    assert ei.value.code == L.ERR_BAD_QUALITY
    slab, qslab, off, ln, *_ = P(b">r1 d\n ANnT \n>r2\nC", b">r1\n 40\t0  -3 12\n>r2\n9\n\n")
    assert list(ln) == [4, 1] and list(slab[:4]) == [40, 0xFF, 0xFE, 12] and slab[16] == 9 and list(qslab[:4]) == [40, 0, 0, 12]
    assert list(P(b">r\nAn\n", b">r\n5 6\n", False)[0][:2]) == [5, 6]
    # parallel path (>= 1 MB): identical rows for any thread count, and the first bad record is the one reported
    rng = np.random.default_rng(5)
    recs = []
    for i in range(5000):
        n = int(rng.integers(1, 500))
        recs.append(("r%d" % i, "".join(rng.choice(list("ACGTN"), size=n)), [int(v) for v in rng.integers(0, 94, size=n)]))
    fa, qu = _fasta_qual_text(recs)
    assert len(fa) > (1 << 20)
    outs = []
    for threads in (1, 5, 16):
        L.lib.moira_set_host_threads(threads)
        outs.append(P(fa, qu, True))
    rows = lambda o, which: [o[which][int(a):int(a) + int(b)] for a, b in zip(o[2], o[3])]
    for other in outs[1:]:
        assert all(np.array_equal(x, y) for x, y in zip(outs[0][3:], other[3:]))
        assert all(np.array_equal(x, y) for x, y in zip(rows(outs[0], 0), rows(other, 0)))
        assert all(np.array_equal(x, y) for x, y in zip(rows(outs[0], 1), rows(other, 1)))
    assert [list(r) for r in rows(outs[2], 1)[:50]] == [[max(v, 0) for v in q] for _, _, q in recs[:50]]
    bad = list(recs)
    bad[3210] = ("r3210", "ACGT", [1, 2, 3])
    fa, qu = _fasta_qual_text(bad)
    for threads in (1, 11):
        L.lib.moira_set_host_threads(threads)
        with pytest.raises(moira_b200.MoiraError) as ei:
            P(fa, qu, True)
        assert "record 3210" in ei.value.message and "LengthMismatchError" in ei.value.message
    L.lib.moira_set_host_threads(0)


def test_line_offsets_cut_paired_texts_at_the_same_records():
    """moira_line_offsets: offsets behind the k-th newline, texts with and without a final newline, queries beyond the end."""
    import moira_b200
    rng = np.random.default_rng(3)
    lines = [bytes(rng.integers(65, 91, int(rng.integers(0, 300))).astype(np.uint8)) for _ in range(5000)]
    for tail in (b"\n", b""):
        text = b"\n".join(lines) + tail
        q = np.array([0, 1, 2, 7, 4000, 4999, 5000, 5001, 99999], np.uint64)
        off, n = moira_b200.line_offsets(text, q)
        assert n == text.count(b"\n")
        pos = [0]
        for i, c in enumerate(text):
            if c == 10:
                pos.append(i + 1)
        want = [pos[int(k)] if int(k) < len(pos) else len(text) for k in q]
        assert off.tolist() == want
    assert moira_b200.line_offsets(b"", np.array([0, 3], np.uint64))[0].tolist() == [0, 0]


def test_index_fastq_equals_parse_fastq_without_a_slab():
    """moira_index_fastq (the record table for flows that hand the text itself to the device) == moira_parse_fastq's byte
    ranges and lengths, with the same errors for malformed records; multi-threaded ranges included."""
    import gzip
    from moira_b200.api import index_fastq
    raw = gzip.open(os.path.join(ROOT, "tests", "golden", "test1.fastq.gz"), "rb").read()
    for text in (raw, raw * 4, raw[:-1], raw + b"@tail\nAC\n+\n", b"", b"\n\n", b"@a\n  ACGT \n+\n IIII\n"):
        a = moira_b200.parse_fastq(text, 33, True)
        b = index_fastq(text)
        assert np.array_equal(a[2], b[0]) and np.array_equal(a[3], b[1]) and np.array_equal(a[4], b[2])
        assert np.array_equal(a[5], b[3]) and np.array_equal(a[6], b[4])
    for bad in (b"@a\nACG\n+\nII\n", b"@a\n\n+\n\n", b"@a\nAC\n+\n\n"):
        with pytest.raises(moira_b200.MoiraError) as e1:
            moira_b200.parse_fastq(bad, 33, True)
        with pytest.raises(moira_b200.MoiraError) as e2:
            index_fastq(bad)
        assert e1.value.code == e2.value.code == L.ERR_PARSE and e1.value.message == e2.value.message
