import gzip
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _ensure_built():
    """Build the checkers (oracle/) and the product library if their binaries are missing."""
    if not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    if not os.path.exists(os.path.join(ROOT, "moira_b200", "libmoira_b200.so")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "moira_b200", "csrc")])


_ensure_built()


@pytest.fixture(scope="session")
def kat():
    return json.load(open(os.path.join(GOLDEN, "kat.json")))


@pytest.fixture(scope="session")
def forward_records():
    from oracle import py_oracle as po
    return po.parse_fastq_text(gzip.open(os.path.join(GOLDEN, "test1.fastq.gz"), "rt").read())


@pytest.fixture(scope="session")
def forward_names():
    return json.load(gzip.open(os.path.join(GOLDEN, "forward_names.json.gz"), "rt"))


@pytest.fixture(scope="session")
def reverse_records():
    import bz2
    from oracle import py_oracle as po
    return po.parse_fastq_text(bz2.open(os.path.join(GOLDEN, "test2.fastq.bz2"), "rt").read())


@pytest.fixture(scope="session")
def paired_names():
    return json.load(gzip.open(os.path.join(GOLDEN, "paired_names.json.gz"), "rt"))


@pytest.fixture(scope="session")
def ref_alignments():
    return json.load(gzip.open(os.path.join(GOLDEN, "ref_alignments.json.gz"), "rt"))


@pytest.fixture(scope="session")
def oracle_contigs(forward_records, reverse_records):
    """(header, contig, quals, overlap, gaps, mismatches) of the 1000 fixture pairs, by the oracle (default arguments)."""
    from oracle import py_oracle as po
    out = []
    for (h, fs, fq), (h2, rs, rq) in zip(forward_records, reverse_records):
        assert h == h2
        out.append((h,) + po.pair_to_contig(fs, fq, rs, rq))
    return out


@pytest.fixture(scope="session")
def contigs():
    return json.load(gzip.open(os.path.join(GOLDEN, "contigs.json.gz"), "rt"))


@pytest.fixture(scope="session")
def ref_outputs():
    return dict(np.load(os.path.join(GOLDEN, "ref_outputs.npz")))


@pytest.fixture(scope="session")
def ctx():
    import moira_b200
    c = moira_b200.Context(0)
    yield c
    c.close()
