#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r02_tests_chernoff.log
timeout 300 python tools/exact_ab.py ccs:2000000 ccs:5000000 2>&1 | tee gpurun_out/r02_chernoff_ab.txt
