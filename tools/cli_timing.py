"""Phase timings of the CLI on a 10 M-read FASTQ in /dev/shm (MOIRA_B200_CLI_TIMING=1): python tools/cli_timing.py [devices]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["MOIRA_B200_CLI_TIMING"] = "1"
import numpy as np, torch
import bench
from moira_b200 import synth, cli
devs = sys.argv[1] if len(sys.argv) > 1 else "0"
slab = synth.generate_v4_device(10_000_000, 20160106, torch.device("cuda", 0)).cpu().numpy()
rec = bench.make_cli_fastq(slab, 1)
path = "/dev/shm/moira_cli_in.fastq"
with open(path, "wb") as fh:
    fh.write(rec)
del rec, slab
for tag, extra in (("collapse_default", []), ("no_collapse_fastq", ["-c", "False", "-o", "fastq"]), ("collapse_default", [])):
    t0 = time.perf_counter()
    rc = cli.main(cli.parse_arguments(["-ffq", path, "-op", "/dev/shm/moira_cli_" + tag, "--devices", devs] + extra), sys.stdout)
    print(tag, "rc", rc, "seconds %.3f" % (time.perf_counter() - t0), flush=True)
    for f in os.listdir("/dev/shm"):
        if f.startswith("moira_cli_" + tag):
            os.remove("/dev/shm/" + f)
os.remove(path)
