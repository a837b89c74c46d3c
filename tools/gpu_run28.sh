#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cli.py tests/test_gpu_contigs.py tests/test_moira_module.py -m gpu -x -q 2>&1 | tail -5
timeout 600 python tools/cli_timing_paired.py 2000000 2>&1 | grep -v "^/dev/shm\|^$\|^The following\|^- " | tee gpurun_out/r02_cli_paired2.txt
