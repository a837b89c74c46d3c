"""The Python-3 host (moira_b200/cli.py) against the reference's full-pipeline goldens
(moira/test/test_moira.py:74-113).  Record ORDER among equal-abundance uniques comes from a
Python-2 dict in the reference (moira.py:492), so files are compared as record sets."""
import bz2
import gzip
import io
import json
import os
import shutil

import pytest

from moira_b200 import cli

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def _records(path, opener=open):
    lines = opener(path, "rt").read().splitlines()
    return {lines[i]: lines[i + 1] for i in range(0, len(lines), 2)}


def _names(path, opener=open):
    out = {}
    for line in opener(path, "rt"):
        rep, members = line.rstrip("\n").split("\t")
        out[rep] = members.split(",")
    return out


def test_argument_surface_matches_reference_defaults():
    a = cli.parse_arguments(["-ffq", "x.fastq"])
    assert (a.alpha, a.uncert, a.maxerrors, a.error_calc, a.ambigs, a.round, a.truncate) == \
        (0.005, 0.01, None, "poisson_binomial", "treat_as_errors", False, None)      # moira.py:649-668
    assert a.collapse is True and a.pipeline == "mothur" and a.output_format == "fasta" and a.fastq_offset == 33
    assert cli.parse_arguments(["-ffq", "x", "-c", "False"]).collapse is False
    with pytest.raises(SystemExit):
        cli.parse_arguments(["-ffq", "x", "-u", "0.02", "-me", "2"])                  # mutually exclusive, moira.py:663-667
    buf = io.StringIO()
    assert cli.check_arguments(cli.parse_arguments(["-ffq", "x", "-a", "1.5"]), buf) is False
    assert "alpha parameter must be between 0" in buf.getvalue()
    buf = io.StringIO()
    assert cli.check_arguments(cli.parse_arguments(["-ffq", "x", "--paired"]), buf) is False        # moira.py:701-705
    assert "You must provide one reverse fastq file" in buf.getvalue()
    a = cli.parse_arguments(["-ffq", "x", "-rfq", "y", "--only_contig"])
    assert cli.check_arguments(a, io.StringIO()) is True and a.paired is True                        # moira.py:695-696
    assert (a.match, a.mismatch, a.gap, a.insert, a.deltaq, a.consensus_qscore, a.qscore_cap, a.trim_overlap) == \
        (1, -1, -2, 20, 6, "best", 40, False)                                                        # moira.py:629-646
    for bad in (["-m", "-1"], ["-x", "1"], ["-g", "1"], ["-i", "0"], ["-d", "0"], ["-mo", "0"]):
        assert cli.check_arguments(cli.parse_arguments(["-ffq", "x", "-rfq", "y", "--paired"] + bad), io.StringIO()) is False
    assert cli.check_arguments(cli.parse_arguments([]), io.StringIO()) is False


def test_open_input_sniffs_compression(tmp_path):
    raw = gzip.open(os.path.join(GOLDEN, "test1.fastq.gz"), "rb").read()
    plain, gz, bz = tmp_path / "a.fastq", tmp_path / "b.anything", tmp_path / "c.anything"
    plain.write_bytes(raw)
    gz.write_bytes(gzip.compress(raw))
    bz.write_bytes(bz2.compress(raw))
    for p in (plain, gz, bz):
        assert cli.open_input(str(p)).read() == raw                                   # moira.py:1058-1090


def test_fastq_text_loading_and_record_aligned_shards(tmp_path, forward_records):
    """The FASTQ flow's host side: load_text (plain = memory-mapped, gz / bz2 inflated) gives the same bytes; the shard
    cuts (one shard per GPU) fall on record starts, so parsing the shards one by one gives the reference's records."""
    import numpy as np
    import moira_b200
    raw = gzip.open(os.path.join(GOLDEN, "test1.fastq.gz"), "rb").read()
    plain, gz, bgzf = tmp_path / "a.fastq", tmp_path / "b.fastq.gz", tmp_path / "c.fastq.gz"
    plain.write_bytes(raw)
    gz.write_bytes(gzip.compress(raw))
    fd = os.open(str(bgzf), os.O_CREAT | os.O_WRONLY, 0o644)             # blocked gzip (bgzip / Illumina writers): inflated in parallel
    moira_b200.gz_deflate(raw, fd, eof=True)
    os.close(fd)
    assert moira_b200.gz_scan(bgzf.read_bytes())[0] > 1 and gzip.open(str(bgzf), "rb").read() == raw
    for path in (plain, gz, bgzf):
        text, _keep = cli.load_text(str(path))
        assert text.tobytes() == raw
    for parts in (1, 2, 3, 7, 8):
        cuts = moira_b200.fastq_split(raw, parts)
        assert cuts[0] == 0 and cuts[-1] == len(raw) and (np.diff(cuts.astype(np.int64)) >= 0).all()
        got = []
        for a, b in zip(cuts[:-1], cuts[1:]):
            piece = raw[int(a):int(b)]
            assert piece[:1] in (b"@", b"")
            slab, off, ln, hoff, hlen, soff, qoff = moira_b200.parse_fastq(piece, 33, True)
            ho2, hl2 = moira_b200.fastq_headers(piece, soff)
            assert np.array_equal(ho2, hoff) and np.array_equal(hl2, hlen)
            got += [(piece[int(h):int(h) + int(l)].decode().replace(":", "_"), piece[int(s_):int(s_) + int(n)].decode())
                    for h, l, s_, n in zip(hoff, hlen, soff, ln)]
        assert got == [(r[0], r[1]) for r in forward_records]


@pytest.mark.gpu
def test_forward_dataset_outputs_match_goldens(tmp_path, forward_names):
    """test_moira.py:74-87 (testProcessForwardDataset) through the new host + CUDA path."""
    fq = tmp_path / "test1.fastq"
    fq.write_bytes(gzip.open(os.path.join(GOLDEN, "test1.fastq.gz"), "rb").read())
    prefix = str(tmp_path / "forward")
    assert cli.run(["-ffq", str(fq), "-op", prefix, "--silent"]) == 0
    golden = json.load(gzip.open(os.path.join(GOLDEN, "forward_outputs.json.gz"), "rt"))
    for lab in ("good", "bad"):
        fa = _records("%s.qc.%s.fasta" % (prefix, lab))
        qu = _records("%s.qc.%s.qual" % (prefix, lab))
        assert set(fa) == set(golden[lab]) == set(qu)
        for hdr, (seq, qual) in golden[lab].items():
            assert fa[hdr] == seq and qu[hdr] == qual
        assert _names("%s.qc.%s.names" % (prefix, lab)) == forward_names[lab]
    assert len(golden["good"]) == 122 and len(golden["bad"]) == 365


@pytest.mark.gpu
def test_compression_and_formats(tmp_path):
    """test_moira.py:102-113 (testCompression) + fastq/USEARCH/no-collapse outputs stay self-consistent."""
    src = os.path.join(GOLDEN, "test1.fastq.gz")
    base = str(tmp_path / "plain")
    assert cli.run(["-ffq", src, "-op", base, "--silent"]) == 0
    plain = open(base + ".qc.good.fasta").read()
    for comp, opener in (("gz", gzip.open), ("bz2", bz2.open)):
        pre = str(tmp_path / comp)
        assert cli.run(["-ffq", src, "-op", pre, "-oc", comp, "--silent"]) == 0
        assert opener("%s.qc.good.fasta.%s" % (pre, comp), "rt").read() == plain
    # blocked gzip in (inflated on all host threads), and the gz outputs above are blocked gzip themselves: one as the next input
    import moira_b200
    bg = str(tmp_path / "in.bgzf.gz")
    fd = os.open(bg, os.O_CREAT | os.O_WRONLY, 0o644)
    moira_b200.gz_deflate(gzip.open(src, "rb").read(), fd, eof=True)
    os.close(fd)
    pre = str(tmp_path / "frombgzf")
    assert cli.run(["-ffq", bg, "-op", pre, "--silent"]) == 0
    assert open(pre + ".qc.good.fasta").read() == plain
    names_gz = open(str(tmp_path / "gz") + ".qc.good.names.gz", "rb").read()
    assert moira_b200.gz_scan(names_gz)[0] >= 1 and moira_b200.gz_inflate(names_gz).tobytes() == open(base + ".qc.good.names", "rb").read()
    # no collapse, fastq output, USEARCH headers, maxerrors mode, truncation
    pre = str(tmp_path / "nc")
    assert cli.run(["-ffq", src, "-op", pre, "-c", "False", "-o", "fastq", "-pi", "USEARCH", "-me", "2", "-t", "200",
                    "--silent"]) == 0
    good = open(pre + ".qc.good.fastq").read().splitlines()
    bad = open(pre + ".qc.bad.fastq").read().splitlines()
    assert (len(good) + len(bad)) // 4 == 1000
    assert all(len(good[i + 1]) == 200 and ";ee=" in good[i] and ";size=1;" in good[i] for i in range(0, len(good), 4))
    assert all("errors > 2.00" in bad[i] or "length below 200" in bad[i] for i in range(0, len(bad), 4))
    for i in range(0, len(good), 4):
        assert float(good[i].split(";ee=")[1].split(";")[0]) <= 2.0
    shutil.rmtree(tmp_path, ignore_errors=True)


@pytest.mark.gpu
def test_fasta_qual_input_reclassifies_golden_contigs(tmp_path, contigs):
    """BASELINE config C1's shape: 253-bp contigs as fasta + qual (the reference's own golden contigs,
    moira/test/test_results/paired.qc.*): 324 good / 76 bad, same labels, same records."""
    fa, qu = tmp_path / "c.fasta", tmp_path / "c.qual"
    with open(fa, "w") as f, open(qu, "w") as q:
        for c in contigs:
            f.write(">%s\textra\n%s\n" % (c["header"], c["seq"]))
            q.write(">%s\n%s\n" % (c["header"], " ".join(map(str, c["quals"]))))
    prefix = str(tmp_path / "out")
    assert cli.run(["-ff", str(fa), "-fq", str(qu), "-op", prefix, "-c", "False", "--silent"]) == 0
    good = _records(prefix + ".qc.good.fasta")
    bad = _records(prefix + ".qc.bad.fasta")
    assert len(good) == 324 and len(bad) == 76
    for c in contigs:
        if c["label"] == "good":
            assert good[">" + c["header"]] == c["seq"]
        else:
            assert bad[">%s\tuncert > 0.010" % c["header"]] == c["seq"]
    with pytest.raises(cli.NameMismatchError):
        bad_q = tmp_path / "bad.qual"
        bad_q.write_text(open(qu).read().replace(contigs[3]["header"], "someone_else", 1))
        cli.run(["-ff", str(fa), "-fq", str(bad_q), "-op", prefix, "--silent"])


@pytest.mark.gpu
def test_paired_dataset_outputs_match_goldens(tmp_path, contigs, paired_names):
    """test_moira.py:88-101 (testProcessPairedDataset) and :102-113 (testCompression: gz forward + bz2 reverse input):
    read pairs -> contigs on the device -> filter -> collapse -> the golden paired.qc.* records and names."""
    prefix = str(tmp_path / "paired")
    assert cli.run(["-ffq", os.path.join(GOLDEN, "test1.fastq.gz"), "-rfq", os.path.join(GOLDEN, "test2.fastq.bz2"), "--paired",
                    "-op", prefix, "--silent"]) == 0
    want = {"good": {}, "bad": {}}
    for c in contigs:
        line = ">%s%s" % (c["header"], "\t" + c["reason"] if c["reason"] else "")
        want[c["label"]][line] = (c["seq"], " ".join(map(str, c["quals"])))
    for lab in ("good", "bad"):
        fa = _records("%s.qc.%s.fasta" % (prefix, lab))
        qu = _records("%s.qc.%s.qual" % (prefix, lab))
        assert set(fa) == set(want[lab]) == set(qu)
        for hdr, (seq, qual) in want[lab].items():
            assert fa[hdr] == seq and qu[hdr] == qual
        assert _names("%s.qc.%s.names" % (prefix, lab)) == paired_names[lab]
    report = open(prefix + ".contigs.report").read().splitlines()
    assert report[0] == "header\tn_seqs\toverlap_length\tgaps\tmismatches" and len(report) == 401
    assert sum(int(line.split("\t")[1]) for line in report[1:]) == 1000


@pytest.mark.gpu
def test_paired_options_only_contig_min_overlap_and_fasta_qual_input(tmp_path, oracle_contigs, forward_records, reverse_records):
    f1, f2 = os.path.join(GOLDEN, "test1.fastq.gz"), os.path.join(GOLDEN, "test2.fastq.bz2")
    by_header = {h: (c, q, ov) for h, c, q, ov, _, _ in oracle_contigs}
    # --only_contig, no collapse: every contig is written to the good files, unfiltered (moira.py:900-908)
    pre = str(tmp_path / "oc")
    assert cli.run(["-ffq", f1, "-rfq", f2, "--only_contig", "-c", "False", "-op", pre, "--silent"]) == 0
    fa, qu = _records(pre + ".qc.good.fasta"), _records(pre + ".qc.good.qual")
    assert len(fa) == 1000 and not open(pre + ".qc.bad.fasta").read()
    for h, (c, q, _) in by_header.items():
        assert fa[">" + h] == c and qu[">" + h] == " ".join(str(max(v, 1)) for v in q)
    # --min_overlap: contigs whose reads overlap by less are rejected with the reference's note (moira.py:886-897)
    cut = sorted(ov for _, _, ov in by_header.values())[100]
    pre = str(tmp_path / "mo")
    assert cli.run(["-ffq", f1, "-rfq", f2, "--paired", "-c", "False", "-mo", str(cut), "-op", pre, "--silent"]) == 0
    bad = _records(pre + ".qc.bad.fasta")
    short = {h for h, (_, _, ov) in by_header.items() if ov < cut}
    assert {k[1:].split("\t")[0] for k in bad if "overlap length below %d" % cut in k} == short and len(short) >= 20
    # the same pairs as fasta + qual files give the same contigs
    names = {}
    for tag, recs in (("f", forward_records), ("r", reverse_records)):
        names[tag] = (str(tmp_path / (tag + ".fasta")), str(tmp_path / (tag + ".qual")))
        with open(names[tag][0], "w") as fa_fh, open(names[tag][1], "w") as qu_fh:
            for h, s_, q in recs[:200]:
                fa_fh.write(">%s\n%s\n" % (h, s_))
                qu_fh.write(">%s\n%s\n" % (h, " ".join(map(str, q))))
    pre = str(tmp_path / "fq")
    assert cli.run(["-ff", names["f"][0], "-fq", names["f"][1], "-rf", names["r"][0], "-rq", names["r"][1], "--only_contig",
                    "-c", "False", "-op", pre, "--silent"]) == 0
    fa = _records(pre + ".qc.good.fasta")
    assert len(fa) == 200
    for h, _, _ in forward_records[:200]:
        assert fa[">" + h] == by_header[h][0]
    # mismatching headers between the two files
    swapped = tmp_path / "swapped.fastq"
    raw = bz2.open(f2, "rb").read().split(b"\n")
    raw[4], raw[0] = raw[0], raw[4]
    swapped.write_bytes(b"\n".join(raw))
    with pytest.raises(cli.NameMismatchError):
        cli.run(["-ffq", f1, "-rfq", str(swapped), "--paired", "-op", str(tmp_path / "x"), "--silent"])
