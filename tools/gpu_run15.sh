#!/bin/bash
# classify-first (exact mode): A/B of the thresholds, then the GPU parity tests with the new default
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 300 python tools/exact_ab.py ccs:2000000 mixed:10000000 2>&1 | sed 's/decision.*| exact/exact/'; }
{
run MOIRA_B200_CLASSIFY_FIRST_K=1000 MOIRA_B200_CLASSIFY_FIRST_SORTED_K=1000
run MOIRA_B200_CLASSIFY_FIRST_K=9 MOIRA_B200_CLASSIFY_FIRST_SORTED_K=3
run MOIRA_B200_CLASSIFY_FIRST_K=1000 MOIRA_B200_CLASSIFY_FIRST_SORTED_K=1000 MOIRA_B200_EXACT_SORTED_KCAP=2
run MOIRA_B200_CLASSIFY_FIRST_K=1000 MOIRA_B200_CLASSIFY_FIRST_SORTED_K=1000 MOIRA_B200_EXACT_SORTED_KCAP=3
run MOIRA_B200_CLASSIFY_FIRST_K=1000 MOIRA_B200_CLASSIFY_FIRST_SORTED_K=1000 MOIRA_B200_EXACT_SORTED_KCAP=4
echo "== classify-first on the fixed-length workloads (k_first 4 / 6)"
MOIRA_B200_CLASSIFY_FIRST_K=3 timeout 300 python tools/exact_ab.py v4:10000000 real:10000000 v3v4:10000000 2>&1 | sed 's/decision.*| exact/exact/'
} > gpurun_out/r02_cf_ab.txt 2>&1
cat gpurun_out/r02_cf_ab.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r02_tests_cf.log
