#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02_tests_j.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r02_tests_j.log
echo "== default"; timeout 300 python tools/exact_ab.py mixed:10000000 v3v4:4000000 2>&1 | tee gpurun_out/r02_exact_ab_sorted16.txt
