#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "capped or multi_stage or classify" 2>&1 | tail -8 | tee gpurun_out/r02_tests_sc.log
run() { echo "== $*"; env "$@" timeout 300 python tools/exact_ab.py $W 2>&1 | sed 's/first pass [0-9.]* ms//g'; }
{
W="mixed:10000000"
run MOIRA_B200_NO_SORTED_CASCADE=1
run MOIRA_B200_NO_SORTED_CASCADE=0
run MOIRA_B200_SORTED_CASCADE_CAP=3
run MOIRA_B200_SORTED_CASCADE_CAP=5
} > gpurun_out/r02_sc_ab.txt 2>&1
cat gpurun_out/r02_sc_ab.txt
