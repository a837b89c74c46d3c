#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r02_tests_cf3.log
timeout 300 python tools/exact_ab.py 2>&1 | tee gpurun_out/r02_cf_ab3.txt
