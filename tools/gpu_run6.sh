#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02_tests_f.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r02_tests_f.log
for v in noroll roll12 roll7 noroll roll12; do
echo "== $v"; MOIRA_B200_LIB=$PWD/build/variants/lib_$v.so timeout 300 python tools/exact_ab.py 2>&1 | tee -a gpurun_out/r02_exact_ab_$v.txt
done
python - <<'PY' > gpurun_out/r02_cli_timing.txt 2>&1
import os, sys, time, io
sys.path.insert(0, os.getcwd())
os.environ["MOIRA_B200_CLI_TIMING"] = "1"
import numpy as np, torch
import bench
from moira_b200 import synth, cli
slab = synth.generate_v4_device(10_000_000, 20160106, torch.device("cuda", 0)).cpu().numpy()
rec = bench.make_cli_fastq(slab, 1)
if isinstance(rec, tuple): rec = rec[0]
path = "/dev/shm/moira_cli_in.fastq"
open(path, "wb").write(rec)
for tag, extra in (("collapse_default", []), ("no_collapse_fastq", ["-c", "False", "-o", "fastq"]), ("collapse_default", [])):
    t0 = time.perf_counter()
    rc = cli.main(cli.parse_arguments(["-ffq", path, "-op", "/dev/shm/moira_cli_" + tag, "--devices", "0"] + extra), sys.stdout)
    print(tag, "rc", rc, "seconds %.3f" % (time.perf_counter() - t0), flush=True)
for f in os.listdir("/dev/shm"):
    if f.startswith("moira_cli_"): os.remove("/dev/shm/" + f)
PY
cat gpurun_out/r02_cli_timing.txt | grep -v "^$" | head -60
