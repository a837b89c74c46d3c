#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r02_tests_cls32.log
timeout 300 python tools/exact_ab.py 2>&1 | tee gpurun_out/r02_cls32_ab.txt
python - <<'PY' 2>&1 | tee gpurun_out/r02_cli_gz.txt
# CLI with compressed files in and out, 4 M reads (tools: the bench's e2e_cli.compressed at a smaller size)
import os, sys, time, io, tempfile, shutil
sys.path.insert(0, os.getcwd())
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
    t = time.time(); n = moira_b200.gz_deflate(rec.reshape(-1), fd, 0, 6, 0, eof=True); print("deflate level 6: %.2f s, %.2f GB/s, ratio %.2f" % (time.time() - t, rec.nbytes / (time.time() - t) / 1e9, rec.nbytes / n)); os.close(fd)
    for tag, args in (("plain in, plain out", ["-ffq", plain]), ("bgzf in, plain out", ["-ffq", gz]), ("bgzf in, gz out", ["-ffq", gz, "-oc", "gz"]), ("plain in, gz out", ["-ffq", plain, "-oc", "gz"])):
        for rep in range(2):
            t = time.time(); rc = cli.main(cli.parse_arguments(args + ["-op", os.path.join(tmp, "o")]), io.StringIO()); dt = time.time() - t
        print("%s: %.2f s = %.3g reads/s rc %d" % (tag, dt, m / dt, rc))
finally:
    shutil.rmtree(tmp, ignore_errors=True)
PY
