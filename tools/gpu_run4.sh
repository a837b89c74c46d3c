#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02_tests_e.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r02_tests_e.log
for gap in 0.5 1.0 2.0; do
echo "== direct rung gap $gap"; MOIRA_B200_DIRECT_GAP=$gap timeout 300 python tools/exact_ab.py 2>&1 | tee gpurun_out/r02_exact_ab_gap$gap.txt | sed 's/decision.*| exact/exact/'
done
echo "== classifier, no multi"; MOIRA_B200_NO_CASCADE_MULTI=1 MOIRA_B200_NO_DIRECT_RUNG=1 timeout 300 python tools/exact_ab.py 2>&1 | tee gpurun_out/r02_exact_ab_classifier.txt
echo "== default"; timeout 300 python tools/exact_ab.py 2>&1 | tee gpurun_out/r02_exact_ab_default.txt | sed 's/| exact.*//'
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ladder_tpr -s 1 -c 1 -o gpurun_out/r02_ladder_v4 python tools/one_step.py 10000000 exact v4 > gpurun_out/r02_ncu_ladder.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_10000000_decision_v3v4.csv python tools/one_step.py 10000000 decision v3v4 > gpurun_out/r02_launches_dec_v3v4.log 2>&1
ls -la gpurun_out/*.ncu-rep
