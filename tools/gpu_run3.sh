#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02_tests_d.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02_tests_d.log
echo "== direct rung"; timeout 300 python tools/exact_ab.py 2>&1 | tee gpurun_out/r02_exact_ab_direct.txt
echo "== classifier"; MOIRA_B200_NO_DIRECT_RUNG=1 timeout 300 python tools/exact_ab.py 2>&1 | tee gpurun_out/r02_exact_ab_classifier.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ladder_tpr -s 1 -c 1 -o gpurun_out/r02_ladder_v4 python tools/one_step.py 10000000 exact v4 > gpurun_out/r02_ncu_ladder.log 2>&1; echo "ncu rc=$?"
