"""Wall clock of the paired CLI (two FASTQ files in /dev/shm -> contigs -> filter -> collapse -> files):
python tools/cli_timing_paired.py [n_pairs] [devices]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["MOIRA_B200_CLI_TIMING"] = "1"
import numpy as np
from tools.bench_contigs import make_pairs
from moira_b200 import cli

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
devs = sys.argv[2] if len(sys.argv) > 2 else "0"
rl = 251
t0 = time.perf_counter()
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
    path = "/dev/shm/moira_pairs_%s.fastq" % tag
    with open(path, "wb") as fh:
        fh.write(rec)
    paths.append(path)
    del rec
print("generated %d pairs in %.1f s" % (n, time.perf_counter() - t0), flush=True)
for tag, extra in (("paired_default", []), ("paired_nocollapse", ["-c", "False"]), ("paired_default", [])):
    t0 = time.perf_counter()
    rc = cli.main(cli.parse_arguments(["-ffq", paths[0], "-rfq", paths[1], "--paired", "-op", "/dev/shm/moira_pairs_" + tag,
                                       "--devices", devs] + extra), sys.stdout)
    dt = time.perf_counter() - t0
    print(tag, "rc", rc, "seconds %.3f -> %.3g pairs/s" % (dt, n / dt), flush=True)
    for f in os.listdir("/dev/shm"):
        if f.startswith("moira_pairs_" + tag):
            os.remove("/dev/shm/" + f)
for p in paths:
    os.remove(p)
