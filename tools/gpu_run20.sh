#!/bin/bash
python - <<'PY' 2>&1 | tee gpurun_out/r02_cli_gz2.txt
import os, sys, time, io, tempfile, shutil
sys.path.insert(0, os.getcwd())
os.environ["MOIRA_B200_CLI_TIMING"] = "1"
import numpy as np
import bench, moira_b200
from moira_b200 import cli, synth
m = 4_000_000
rows, off, ln = synth.generate("v4", m, 5)
rec = bench.make_cli_fastq(rows.reshape(m, -1), 7)
tmp = tempfile.mkdtemp(dir="/dev/shm")
try:
    plain = os.path.join(tmp, "in.fastq"); open(plain, "wb").write(rec)
    gz = os.path.join(tmp, "in.gz"); fd = os.open(gz, os.O_CREAT | os.O_WRONLY)
    n = moira_b200.gz_deflate(rec.reshape(-1), fd, 0, 1, 0, eof=True); os.close(fd)
    del rec, rows
    for tag, args in (("plain in, plain out", ["-ffq", plain]), ("bgzf in, plain out", ["-ffq", gz]), ("bgzf in, gz out", ["-ffq", gz, "-oc", "gz"])):
        for rep in range(2):
            log = io.StringIO()
            t = time.time(); rc = cli.main(cli.parse_arguments(args + ["-op", os.path.join(tmp, "o"), "--devices", "0"]), log); dt = time.time() - t
        print("%s: %.2f s = %.3g reads/s rc %d" % (tag, dt, m / dt, rc))
        print("\n".join(l for l in log.getvalue().splitlines() if "timing" in l or "processed" in l))
finally:
    shutil.rmtree(tmp, ignore_errors=True)
PY
