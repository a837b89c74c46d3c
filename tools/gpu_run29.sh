#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cli.py tests/test_gz.py tests/test_moira_module.py -x -q 2>&1 | tail -4
python - <<'PY' 2>&1 | tee gpurun_out/r02_cli_gz3.txt
# CLI from compressed inputs with the library's own DEFLATE decoder and with zlib only (4 M reads)
import os, sys, time, io, tempfile, shutil, zlib, subprocess
sys.path.insert(0, os.getcwd())
import numpy as np
import bench, moira_b200
from moira_b200 import cli, synth
m = 4_000_000
rows, off, ln = synth.generate("v4", m, 5)
rec = bench.make_cli_fastq(rows.reshape(m, -1), 7)
tmp = tempfile.mkdtemp(dir="/dev/shm")
try:
    gz = os.path.join(tmp, "in.bgzf.gz"); fd = os.open(gz, os.O_CREAT | os.O_WRONLY)
    moira_b200.gz_deflate(rec.reshape(-1), fd, 0, 6, 0, eof=True); os.close(fd)
    plain = os.path.join(tmp, "in.plain.gz")
    zc = zlib.compressobj(6, zlib.DEFLATED, 31)
    with open(plain, "wb") as fh:
        fh.write(zc.compress(rec[:1_000_000].tobytes())); fh.write(zc.flush())
    nbytes = rec.nbytes; pbytes = rec[:1_000_000].nbytes
    del rec, rows
    code = ("import sys, time, numpy as np; sys.path.insert(0, %r); import moira_b200\n"
            "z = np.fromfile(sys.argv[1], dtype=np.uint8)\n"
            "for _ in range(2):\n    t = time.time(); out = moira_b200.gz_inflate(z); dt = time.time() - t\n"
            "print('%%.2f s  %%.2f GB/s' %% (dt, out.nbytes / dt / 1e9))") % os.getcwd()
    for env in ("0", "1"):
        for name, path in (("blocked gzip, all threads", gz), ("one plain member (1 M reads), one thread", plain)):
            r = subprocess.run([sys.executable, "-c", code, path], env=dict(os.environ, MOIRA_B200_ZLIB_INFLATE=env), capture_output=True, text=True)
            print("%s | %s: %s" % ("zlib only" if env == "1" else "own decoder", name, (r.stdout + r.stderr).strip()))
    for tag, args in (("bgzf in, plain out", ["-ffq", gz]), ("plain gzip in (1 M reads), plain out", ["-ffq", plain])):
        for rep in range(2):
            t = time.time(); rc = cli.main(cli.parse_arguments(args + ["-op", os.path.join(tmp, "o"), "--devices", "0"]), io.StringIO()); dt = time.time() - t
        print("%s: %.2f s rc %d" % (tag, dt, rc))
finally:
    shutil.rmtree(tmp, ignore_errors=True)
PY
