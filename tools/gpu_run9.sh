#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02_tests_i.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r02_tests_i.log
echo "== default"; timeout 300 python tools/exact_ab.py 2>&1 | tee gpurun_out/r02_exact_ab_default2.txt
echo "== no multi"; MOIRA_B200_NO_CASCADE_MULTI=1 timeout 300 python tools/exact_ab.py v3v4:10000000 mixed:10000000 2>&1 | tee gpurun_out/r02_exact_ab_nomulti2.txt
