#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02_tests_e.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r02_tests_e.log
echo "== default"; timeout 300 python tools/exact_ab.py 2>&1 | tee gpurun_out/r02_exact_ab_default.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ladder_tpr -s 1 -c 1 -o /tmp/r02_ladder_v4 python tools/one_step.py 10000000 exact v4 > gpurun_out/r02_ncu_ladder.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/r02_ladder_v4.ncu-rep --page raw --csv > gpurun_out/r02_ladder_v4_raw.csv 2>/dev/null
ncu -i /tmp/r02_ladder_v4.ncu-rep --page source --csv > gpurun_out/r02_ladder_v4_source.csv 2>/dev/null
ls -la gpurun_out/r02_ladder*
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_10000000_decision_v3v4.csv python tools/one_step.py 10000000 decision v3v4 > gpurun_out/r02_launches_dec_v3v4.log 2>&1
