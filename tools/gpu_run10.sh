#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02_tests_k.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r02_tests_k.log
echo "== default"; timeout 300 python tools/exact_ab.py 2>&1 | tee gpurun_out/r02_exact_ab_ladder16.txt
