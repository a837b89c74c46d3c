#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "every_rule" 2>&1 | tail -15 | tee gpurun_out/r02_tests_rules.log
