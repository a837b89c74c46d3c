#!/usr/bin/env python3
"""Generate the committed golden fixtures under tests/golden/ from the reference tree.

Run in the build container only (needs /root/reference and oracle/_ref/bernoulli.so, i.e.
`make -C oracle` first):   python tests/golden/make_golden.py

Outputs (all small, committed):
  kat.json               the reference's own known-answer vectors (moira/test/test_moira.py:39-45,
                         63-70, 118-128) -- inputs and expected values copied verbatim as DATA
  test1.fastq.gz         the reference's 1000-read MiSeq forward fixture (moira/test/test1.fastq.gz)
  forward_names.json.gz  golden .names partitions of the forward full-pipeline test
                         (moira/test/test_results/forward.qc.{good,bad}.names)
  forward_outputs.json.gz  golden fasta/qual records of the same test (header line -> [sequence, quality line])
  contigs.json.gz        the 400 golden paired contigs (fasta+qual) with good/bad labels
                         (moira/test/test_results/paired.qc.{good,bad}.{fasta,qual})
  test2.fastq.bz2        the matching reverse reads (moira/test/test2.fastq.bz2), for the paired pipeline
  paired_names.json.gz   golden .names partitions of the paired full-pipeline test
                         (moira/test/test_results/paired.qc.{good,bad}.names)
  ref_alignments.json.gz outputs of the UNMODIFIED reference aligner (oracle/_ref/nw_align.so, cythonised
                         from moira/nw_align.pyx) on 240 seeded pairs, several scoring schemes
  ref_outputs.npz        outputs of the UNMODIFIED reference binary (oracle/_ref) on: the 1000
                         forward reads, the 400 contigs, and 4000 seeded synthetic reads covering
                         N/n, Q=0, L=1..600, several alphas -- as packed slabs + (ee, Ns)
"""
import ast
import gzip
import json
import os
import re
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import py_oracle as po  # noqa: E402

REF = "/root/reference/moira"
T = open(os.path.join(REF, "test", "test_moira.py")).read()


def grab(name):
    m = re.search(r"^%s = (.*)$" % re.escape(name), T, re.M)
    return ast.literal_eval(m.group(1))


def _dump_gz(name, obj):
    """json -> gzip with a fixed mtime so regenerating the fixtures is byte-reproducible."""
    with open(os.path.join(HERE, name), "wb") as raw:
        with gzip.GzipFile(fileobj=raw, mode="wb", mtime=0) as gz:
            gz.write(json.dumps(obj).encode())


def main():
    ref = po.ref_module()

    # ---- known-answer vectors ---------------------------------------------------------------
    s1, s2 = grab("testSeq1"), grab("testSeq2")
    q1s, q2s = grab("testQual1"), grab("testQual2")
    fwd = grab("test_ForwardProcess")
    par = grab("test_PairedProcess")
    pb_expect = re.search(r"calculate_errors_PB\(testSeq1, testQual1, 0.005\), \(([0-9.]+), 0\)", T).group(1)
    po_expect = re.search(r"calculate_errors_poisson\(testSeq1, testQual1, 0.005\), \(([0-9.]+), 0\)", T).group(1)
    kat = {
        "source": "moira/test/test_moira.py",
        "alpha": 0.005, "fastq_offset": 33,
        "testSeq1": s1, "testQual1_ascii": q1s, "testSeq2": s2, "testQual2_ascii": q2s,
        "pb_expected": [float(pb_expect), 0],
        "poisson_expected": [float(po_expect), 0],
        "forward_process": {"truncate": 200, "seq": fwd[1], "quals": list(fwd[2]), "ee": fwd[3]},
        "paired_process": {"seq": par[1], "quals": list(par[2]), "ee": par[3],
                           "overlap": par[4], "gaps": par[5], "mismatches": par[6]},
    }
    # contig constructor vectors (test_moira.py:49-59, 118-135) and the arguments they were made with
    rc2 = grab("testRC2")
    al = grab("test_aligned")
    ct = grab("test_contig")
    kat["contig_args"] = {"match": 1, "mismatch": -1, "gap": -2, "insert": 20, "deltaq": 6, "consensus_qscore": "best",
                          "qscore_cap": 40, "trim_overlap": False}
    kat["testRC2"] = [rc2[0], list(rc2[1])]
    kat["test_aligned"] = [al[0], al[1], al[2]]
    kat["test_contig"] = [ct[0], list(ct[1]), ct[2], ct[3], ct[4]]
    q1 = [ord(c) - 33 for c in q1s]
    assert ref.calculate_errors_PB(s1, q1, 0.005) == tuple(kat["pb_expected"])
    json.dump(kat, open(os.path.join(HERE, "kat.json"), "w"), indent=1)

    # ---- forward fastq fixture + golden names -------------------------------------------------
    shutil.copyfile(os.path.join(REF, "test", "test1.fastq.gz"), os.path.join(HERE, "test1.fastq.gz"))
    names = {}
    for lab in ("good", "bad"):
        d = {}
        for line in open(os.path.join(REF, "test", "test_results", "forward.qc.%s.names" % lab)):
            rep, members = line.rstrip("\n").split("\t")
            d[rep] = members.split(",")
        names[lab] = d
    _dump_gz("forward_names.json.gz", names)
    shutil.copyfile(os.path.join(REF, "test", "test2.fastq.bz2"), os.path.join(HERE, "test2.fastq.bz2"))
    names = {}
    for lab in ("good", "bad"):
        d = {}
        for line in open(os.path.join(REF, "test", "test_results", "paired.qc.%s.names" % lab)):
            rep, members = line.rstrip("\n").split("\t")
            d[rep] = members.split(",")
        names[lab] = d
    _dump_gz("paired_names.json.gz", names)

    # ---- the unmodified reference aligner on seeded pairs -------------------------------------
    nw = po.ref_nw_module()
    rng = np.random.Generator(np.random.PCG64(20160401))
    schemes = [(1, -1, -2), (2, -3, -5), (1, -1, -1), (0, 0, 0), (5, -4, -10), (1, -2, 0)]
    cases = []
    for it in range(240):
        l1, l2 = int(rng.integers(1, 140)), int(rng.integers(1, 140))
        a = "".join(rng.choice(list("ACGT"), l1))
        if it % 3:      # overlapping pair with a few substitutions / indels, like a read pair
            k = int(rng.integers(0, l1))
            b = list(a[k:] + "".join(rng.choice(list("ACGT"), l2)))
            b = [c if rng.random() > 0.06 else "ACGTN"[int(rng.integers(5))] for c in b]
            if rng.random() < 0.5 and len(b) > 4:
                del b[int(rng.integers(len(b)))]
            b = "".join(b)[:max(1, l2)]
        else:
            b = "".join(rng.choice(list("ACGTN"), l2))
        m, x, g = schemes[it % len(schemes)]
        a1, a2, score = nw.nw_align(a, b, m, x, g)
        cases.append({"seq_1": a, "seq_2": b, "match": m, "mismatch": x, "gap": g, "aligned_1": a1, "aligned_2": a2,
                      "score": int(score)})
    _dump_gz("ref_alignments.json.gz", cases)

    # ---- golden forward output records (fasta + qual, keyed by header; order is Py2-dict dependent) ---
    fwd_out = {}
    for lab in ("good", "bad"):
        fa = open(os.path.join(REF, "test", "test_results", "forward.qc.%s.fasta" % lab)).read().splitlines()
        qu = open(os.path.join(REF, "test", "test_results", "forward.qc.%s.qual" % lab)).read().splitlines()
        d = {}
        for i in range(0, len(fa), 2):
            assert fa[i] == qu[i]
            d[fa[i]] = [fa[i + 1], qu[i + 1]]
        fwd_out[lab] = d
    _dump_gz("forward_outputs.json.gz", fwd_out)

    # ---- golden contigs (paired run outputs, used as single-end fasta+qual inputs) ------------
    contigs = []
    for lab in ("good", "bad"):
        fa = open(os.path.join(REF, "test", "test_results", "paired.qc.%s.fasta" % lab)).read()
        qu = open(os.path.join(REF, "test", "test_results", "paired.qc.%s.qual" % lab)).read()
        hdr_reason = {}
        for line in fa.splitlines()[0::2]:
            parts = line[1:].split("\t")
            hdr_reason[parts[0]] = parts[1] if len(parts) > 1 else ""
        for header, seq, quals in po.parse_fasta_qual_text(fa, qu):
            contigs.append({"header": header, "seq": seq, "quals": quals, "label": lab,
                            "reason": hdr_reason.get(header, "")})
    _dump_gz("contigs.json.gz", contigs)

    # ---- reference-binary outputs -------------------------------------------------------------
    out = {}
    fq = gzip.open(os.path.join(HERE, "test1.fastq.gz"), "rt").read()
    recs = po.parse_fastq_text(fq)
    ee = np.array([ref.calculate_errors_PB(s, [q if q > 0 else 1 for q in ql], 0.005)[0] for _, s, ql in recs])
    out["forward_ee"] = ee
    cee = np.array([ref.calculate_errors_PB(c["seq"], c["quals"], 0.005)[0] for c in contigs])
    out["contigs_ee"] = cee

    rng = np.random.Generator(np.random.PCG64(20160105))
    seqs, quals, alphas = [], [], []
    alpha_choices = [0.005, 0.001, 0.05, 0.5, 0.2, 1e-6]
    for i in range(4000):
        kind = i % 8
        if kind == 0:
            L = int(rng.integers(1, 17))
        elif kind == 1:
            L = int(rng.integers(1, 601))
        else:
            L = int(rng.integers(100, 460))
        if kind in (2, 3):          # clean Illumina-like
            q = rng.choice([40, 39, 38, 37, 36, 35, 33, 30, 25, 20, 12, 2], size=L,
                           p=[.12, .32, .30, .06, .07, .05, .03, .02, .01, .01, .005, .005])
        elif kind == 4:             # very noisy
            q = rng.integers(0, 20, size=L)
        elif kind == 5:             # full uint8-safe range incl. 0 and big scores
            q = rng.integers(0, 94, size=L)
        else:
            q = rng.integers(2, 42, size=L)
        bases = rng.choice(list("ACGT"), size=L)
        pn = [0.0, 0.02, 0.0, 0.01, 0.05, 0.0, 0.3, 1.0][kind] if kind != 7 or i % 16 == 7 else 0.0
        mask = rng.random(L) < pn
        bases = np.where(mask, np.where(rng.random(L) < 0.3, "n", "N"), bases)
        seqs.append("".join(bases.tolist()))
        quals.append([int(v) for v in q])
        alphas.append(alpha_choices[(i // 8) % len(alpha_choices)] if kind != 2 else 0.005)
    syn_ee = np.zeros(len(seqs))
    syn_ns = np.zeros(len(seqs), dtype=np.int32)
    for i, (s, q, a) in enumerate(zip(seqs, quals, alphas)):
        syn_ee[i], syn_ns[i] = ref.calculate_errors_PB(s, q, a)   # binding maps Q==0 -> 1 itself
    assert not np.isnan(syn_ee).any()
    slab, offsets, lengths = po.pack_records(seqs, quals, lower_n_ambiguous=True)
    out.update(syn_slab=slab, syn_offsets=offsets, syn_lengths=lengths,
               syn_alpha=np.array(alphas), syn_ee=syn_ee, syn_ns=syn_ns)
    np.savez_compressed(os.path.join(HERE, "ref_outputs.npz"), **out)
    for f in sorted(os.listdir(HERE)):
        print("%10d  %s" % (os.path.getsize(os.path.join(HERE, f)), f))


if __name__ == "__main__":
    main()
