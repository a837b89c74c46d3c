#!/bin/bash
# classify-first in decision mode (ccs), K cap of the length-bucketed exact first pass (mixed), then the GPU tests
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 300 python tools/exact_ab.py $W 2>&1; }
{
W="ccs:2000000"
run MOIRA_B200_CLASSIFY_FIRST_K=1000 MOIRA_B200_CLASSIFY_FIRST_DEC_K=1000
run MOIRA_B200_CLASSIFY_FIRST_K=9 MOIRA_B200_CLASSIFY_FIRST_DEC_K=9
W="mixed:10000000"
run MOIRA_B200_EXACT_SORTED_KCAP=4
run MOIRA_B200_EXACT_SORTED_KCAP=5
run MOIRA_B200_EXACT_SORTED_KCAP=6
W="v4:2000000 real:2000000 v3v4:2000000 mixed:2000000 ccs:500000"
run MOIRA_B200_EXACT_SORTED_KCAP=4
} > gpurun_out/r02_cf_ab2.txt 2>&1
cat gpurun_out/r02_cf_ab2.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r02_tests_cf2.log
