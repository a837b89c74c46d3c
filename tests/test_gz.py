"""gzip inputs and outputs on the host threads (moira_gz.cpp; reference: gzip.GzipFile behind moira.py:1065-1068 and the
--output_compression gz files of moira.py:323-370).  Host functions: no GPU needed.  The checker is Python's gzip module."""
import gzip
import os
import zlib

import numpy as np
import pytest

import moira_b200
from moira_b200 import _lib as L
from moira_b200.api import gz_deflate, gz_inflate, gz_scan

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _texts():
    rng = np.random.default_rng(4)
    fastq = gzip.open(os.path.join(GOLDEN, "test1.fastq.gz"), "rb").read()
    return {
        "empty": b"",
        "one byte": b"x",
        "fastq": fastq,
        "one piece exactly": bytes(0xff00),
        "one piece + 1": bytes(0xff00 + 1),
        "incompressible": rng.integers(0, 256, 700_001, dtype=np.uint8).tobytes(),     # members that have to be stored
        "mixed": fastq[:300_000] + rng.integers(0, 256, 200_000, dtype=np.uint8).tobytes() + fastq[:123_457],
    }


def _bgzf(data, tmp_path, level=6, threads=0, eof=True):
    fn = str(tmp_path / "t.gz")
    fd = os.open(fn, os.O_CREAT | os.O_TRUNC | os.O_WRONLY, 0o644)
    try:
        n = gz_deflate(data, fd, 0, level, threads, eof=eof)
    finally:
        os.close(fd)
    raw = open(fn, "rb").read()
    assert len(raw) == n
    return raw


@pytest.mark.parametrize("name", list(_texts()))
def test_bgzf_written_here_is_gzip_and_reads_back_in_parallel(tmp_path, name):
    data = _texts()[name]
    for level, threads, eof in ((6, 0, True), (1, 3, False), (9, 1, True)):
        raw = _bgzf(data, tmp_path, level, threads, eof)
        assert gzip.decompress(raw) == data if raw else data == b""                   # any gunzip reads it
        members, total = gz_scan(raw) if raw else (0, 0)
        assert total == len(data)
        assert members == (len(data) + 0xff00 - 1) // 0xff00 + (1 if eof else 0)        # 65 280-byte pieces (+ the empty end member)
        if raw:
            for t in (0, 1, 5):
                assert gz_inflate(raw, t).tobytes() == data
    # members are independent: any run of whole members is a gzip file of its own
    if len(data) > 3 * 0xff00:
        raw = _bgzf(data, tmp_path, eof=False)
        at, sizes = 0, []
        while at < len(raw):
            sizes.append(int.from_bytes(raw[at + 16:at + 18], "little") + 1)
            at += sizes[-1]
        cut = sum(sizes[:2])
        assert gz_inflate(raw[cut:]).tobytes() == data[2 * 0xff00:]


@pytest.mark.parametrize("name", ["empty", "fastq", "incompressible", "mixed"])
def test_plain_gzip_files_one_thread_like_the_gzip_module(name):
    data = _texts()[name]
    z = gzip.compress(data)
    assert gz_scan(z) == (0, 0)                                                         # nothing to cut at
    assert gz_inflate(z).tobytes() == data
    # concatenated members and zero padding behind the last one (gzip.GzipFile reads both)
    both = z + gzip.compress(data[::-1], 1) + b"\0" * 37
    assert gzip.decompress(both) == data + data[::-1]
    assert gz_inflate(both).tobytes() == data + data[::-1]
    # a header with a file name (FNAME), as `gzip file` writes it
    named = bytearray(z[:10]) + b"name.fastq\0" + z[10:]
    named[3] |= 8
    assert gzip.decompress(bytes(named)) == data and gz_inflate(bytes(named)).tobytes() == data


def test_mixed_bgzf_and_plain_members(tmp_path):
    data = _texts()["fastq"]
    raw = _bgzf(data, tmp_path, eof=False) + gzip.compress(b"tail")
    assert gz_scan(raw) == (0, 0)                                                       # one member without a size field: the stream route
    assert gz_inflate(raw).tobytes() == data + b"tail" == gzip.decompress(raw)


def test_bad_gzip_inputs_fail_like_the_gzip_module(tmp_path):
    data = _texts()["fastq"]
    z, raw = gzip.compress(data), _bgzf(data, tmp_path)
    flipped = bytearray(raw)
    flipped[len(raw) // 2] ^= 0x55
    wrong_crc = bytearray(z)
    wrong_crc[-8] ^= 1
    for bad in (b"plain text", z[:-9], z[:200], z + b"junk", raw[:-45], bytes(flipped), bytes(wrong_crc), b"\x1f"):
        with pytest.raises((OSError, EOFError, zlib.error)):
            gzip.decompress(bad)
        with pytest.raises(moira_b200.MoiraError) as err:
            gz_inflate(bad)
        assert err.value.code == L.ERR_PARSE


def test_large_stream_grows_its_buffer():
    """A single member that inflates to much more than four times its size (the first guess of the output buffer)."""
    data = bytes(200 << 20)
    z = zlib.compressobj(6, zlib.DEFLATED, 31)
    blob = z.compress(data) + z.flush()
    out = gz_inflate(blob)
    assert out.size == len(data) and not out.any()


def test_own_decoder_against_zlib_on_many_streams_and_corruptions(tmp_path):
    """The library's DEFLATE decoder (moira_inflate.h) sits in front of zlib: every kind of block (stored, fixed, dynamic; levels
    0..9; long matches; empty members), then 300 random truncations / bit flips / garbage runs of plain and blocked gzip files --
    the result is Python's gzip result, error for error (a decoder that disagrees with a member's CRC-32 hands over to zlib)."""
    rng = np.random.default_rng(6)
    fastq = gzip.open(os.path.join(GOLDEN, "test1.fastq.gz"), "rb").read()
    texts = [b"", b"a", b"ab" * 70000, bytes(100000), fastq, rng.integers(0, 256, 200000, dtype=np.uint8).tobytes(),
             fastq[:50000] + bytes(rng.integers(0, 4, 300000, dtype=np.uint8) + 65) + fastq[:1000] * 40]
    for data in texts:
        for level in (0, 1, 4, 6, 9):
            z = gzip.compress(data, level)
            assert gz_inflate(z).tobytes() == data
            co = zlib.compressobj(level, zlib.DEFLATED, 31, 9, zlib.Z_FIXED)          # fixed Huffman blocks
            z = co.compress(data) + co.flush()
            assert gz_inflate(z).tobytes() == data
            co = zlib.compressobj(level, zlib.DEFLATED, 31)                             # many small blocks: full flushes
            z = b"".join(co.compress(data[i:i + 7001]) + co.flush(zlib.Z_FULL_FLUSH) for i in range(0, len(data), 7001)) + co.flush()
            assert gz_inflate(z).tobytes() == data
    plain, blocked = gzip.compress(fastq, 6), _bgzf(fastq, tmp_path)
    for trial in range(300):
        src = bytearray(plain if trial % 2 else blocked)
        kind = trial % 3
        if kind == 0:
            src = src[:int(rng.integers(1, len(src)))]
        elif kind == 1:
            for _ in range(int(rng.integers(1, 4))):
                src[int(rng.integers(0, len(src)))] ^= 1 << int(rng.integers(0, 8))
        else:
            a = int(rng.integers(0, len(src) - 64))
            src[a:a + 64] = rng.integers(0, 256, 64, dtype=np.uint8).tobytes()
        src = bytes(src)
        try:
            want = gzip.decompress(src)
        except Exception:
            want = None
        try:
            got = gz_inflate(src).tobytes()
        except moira_b200.MoiraError as exc:
            assert exc.code == L.ERR_PARSE
            got = None
        assert got == want, (trial, kind)


def test_own_compressor_level_1_is_read_by_every_inflater(tmp_path):
    """Level 1 = the library's own DEFLATE compressor (moira_deflate.h; each piece inflated and compared before it is written):
    Python's gzip and zlib, and the library's two readers, all get the input back -- text, runs, noise (stored members), tiny
    pieces, piece-boundary sizes."""
    rng = np.random.default_rng(9)
    fastq = gzip.open(os.path.join(GOLDEN, "test1.fastq.gz"), "rb").read()
    texts = [b"", b"x", b"abc", b"aaaa" * 5, fastq, fastq[:0xff00], fastq[:0xff00 + 1], bytes(300000), b"ab" * 100001,
             rng.integers(0, 256, 150000, dtype=np.uint8).tobytes(), bytes(rng.integers(65, 69, 400000, dtype=np.uint8)),
             fastq[:70000] + rng.integers(0, 256, 70000, dtype=np.uint8).tobytes() + fastq[:70000]]
    for data in texts:
        raw = _bgzf(data, tmp_path, level=1, threads=3, eof=True)
        assert gzip.decompress(raw) == data
        assert gz_inflate(raw).tobytes() == data and gz_inflate(raw, 1).tobytes() == data
        # member by member through zlib's raw inflate (what htslib does)
        at, got = 0, []
        while at < len(raw):
            size = int.from_bytes(raw[at + 16:at + 18], "little") + 1
            got.append(zlib.decompress(raw[at + 18:at + size - 8], -15))
            assert zlib.crc32(got[-1]) == int.from_bytes(raw[at + size - 8:at + size - 4], "little")
            at += size
        assert b"".join(got) == data
    # it shrinks text at least as well as zlib's level 1
    z1 = zlib.compressobj(1, zlib.DEFLATED, -15)
    assert len(_bgzf(fastq, tmp_path, level=1, eof=False)) < 1.05 * len(z1.compress(fastq) + z1.flush())
