#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_moira_module.py tests/test_cli.py tests/test_gpu_contigs.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r02_tests_module.log
