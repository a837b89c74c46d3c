"""Wall clock of the FASTA + QUAL CLI (config C1's format): python tools/cli_timing_fasta.py [n_reads]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["MOIRA_B200_CLI_TIMING"] = "1"
import numpy as np
from moira_b200 import synth, cli

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
slab, off, ln = synth.generate("v4", n, 20160105)
q = slab.reshape(n, 256)[:, :253]
isn = q == 0xFF
rng = np.random.default_rng(1)
bases = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, (n, 253))]
bases = np.where(isn, ord("N"), bases).astype(np.uint8)
ids = np.char.zfill(np.arange(n).astype("U8"), 8)
idb = np.frombuffer("".join(ids.tolist()).encode(), dtype=np.uint8).reshape(n, 8)
fa = np.empty((n, 1 + 1 + 8 + 1 + 253 + 1), np.uint8)
fa[:, 0] = ord(">"); fa[:, 1] = ord("r"); fa[:, 2:10] = idb; fa[:, 10] = 10; fa[:, 11:264] = bases; fa[:, 264] = 10
qv = np.where(isn, 2, q).astype(np.uint8)
# two-digit qualities (pad single digits with a leading space, which int() ignores)
tens, ones = qv // 10, qv % 10
qt = np.empty((n, 253, 3), np.uint8)
qt[:, :, 0] = np.where(tens > 0, tens + 48, 32); qt[:, :, 1] = ones + 48; qt[:, :, 2] = 32
qu = np.empty((n, 11 + 253 * 3), np.uint8)
qu[:, 0] = ord(">"); qu[:, 1] = ord("r"); qu[:, 2:10] = idb; qu[:, 10] = 10; qu[:, 11:] = qt.reshape(n, -1); qu[:, -1] = 10
paths = ["/dev/shm/moira_fa.fasta", "/dev/shm/moira_fa.qual"]
open(paths[0], "wb").write(fa); open(paths[1], "wb").write(qu)
print("files: %.2f + %.2f GB" % (fa.nbytes / 1e9, qu.nbytes / 1e9), flush=True)
del fa, qu, qt
for tag, extra in (("fq_default", []), ("fq_nocollapse", ["-c", "False"]), ("fq_default", [])):
    t0 = time.perf_counter()
    rc = cli.main(cli.parse_arguments(["-ff", paths[0], "-fq", paths[1], "-op", "/dev/shm/moira_fa_" + tag] + extra), sys.stdout)
    dt = time.perf_counter() - t0
    print(tag, "rc", rc, "seconds %.3f -> %.3g reads/s" % (dt, n / dt), flush=True)
    for f in os.listdir("/dev/shm"):
        if f.startswith("moira_fa_" + tag):
            os.remove("/dev/shm/" + f)
for p in paths:
    os.remove(p)
