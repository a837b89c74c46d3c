"""Host side of the paired flow alone: python tools/pair_reader_timing.py [n_pairs]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tools.bench_contigs import make_pairs
from moira_b200 import cli, Context, ContigParams, FilterParams

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
rl = 251
fwd, rev = make_pairs(n, rl)
paths = []
for tag, (bases, quals, off, ln) in (("R1", fwd), ("R2", rev)):
    rec = np.empty((n, 10 + 1 + rl + 3 + rl + 1), dtype=np.uint8)
    ids = np.char.zfill(np.arange(n).astype("U8"), 8)
    rec[:, 0] = ord("@"); rec[:, 1] = ord("p")
    rec[:, 2:10] = np.frombuffer("".join(ids.tolist()).encode(), dtype=np.uint8).reshape(n, 8)
    rec[:, 10] = 10
    rec[:, 11:11 + rl] = bases.reshape(n, rl)
    rec[:, 11 + rl:14 + rl] = np.frombuffer(b"\n+\n", dtype=np.uint8)
    rec[:, 14 + rl:14 + 2 * rl] = quals.reshape(n, rl) + 33
    rec[:, -1] = 10
    path = "/dev/shm/moira_prt_%s.fastq" % tag
    with open(path, "wb") as fh:
        fh.write(rec)
    paths.append(path)
args = cli.parse_arguments(["-ffq", paths[0], "-rfq", paths[1], "--paired", "-op", "/dev/shm/moira_prt_out"])
for rep in range(2):
    t0 = time.perf_counter()
    blocks = list(cli._read_fastq_pair_batches(args, True))
    print("reader alone: %.3f s for %d pairs in %d blocks" % (time.perf_counter() - t0, n, len(blocks)), flush=True)
ctx = Context(0)
p = FilterParams(exact_ee=True, ee_output="final")
for rep in range(2):
    t0 = time.perf_counter()
    for ftext, hoff, hlen, fw, rv in blocks:
        ctx.filter_pairs(fw[0], fw[1], fw[2], fw[4], rv[0], rv[1], rv[2], rv[4], ContigParams(), p, True, fw[3], rv[3], fw[5])
    print("filter_pairs alone (pageable in/out): %.3f s" % (time.perf_counter() - t0), flush=True)
ctx.close()
for pth in paths:
    os.remove(pth)
