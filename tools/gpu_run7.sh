#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02_tests_g.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r02_tests_g.log
for v in lroll1 lroll2 lroll1k8 lroll1 lroll2 lroll1k8; do
echo "== $v"; MOIRA_B200_LIB=$PWD/build/variants/lib_$v.so timeout 300 python tools/exact_ab.py 2>&1 | tee -a gpurun_out/r02_exact_ab_$v.txt | sed 's/decision.*| exact/exact/'
done
python tools/cli_timing.py > gpurun_out/r02_cli_timing.txt 2>&1
grep "timing\|seconds\|processed" gpurun_out/r02_cli_timing.txt
