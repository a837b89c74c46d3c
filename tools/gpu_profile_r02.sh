#!/bin/bash
# round-2 evidence: launch list of the bench command + full ncu captures of the two first-pass kernels and the ladder,
# exported as CSV (the .ncu-rep files stay on the box: too large for gpurun_out)
mkdir -p gpurun_out
if [ "$1" != "captures" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tpr_kernel|ladder|wpr_|blk_|policy|count_marks|sorted_first|len_|grp_|dedup|seq_hash|fp64_peak|unpack|Device|fq_|contig" -c 3000 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-cli > gpurun_out/r02_launches.log 2>&1; echo "launch list rc=$?"
fi
cap() {  # tag kernel-regex skip mode profile n
  timeout 900 ncu --set full --clock-control none --import-source on -k $2 -s $3 -c 1 -o /tmp/$1 python tools/one_step.py $6 $4 $5 > gpurun_out/$1.log 2>&1
  echo "ncu $1 rc=$?"
  ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
  ncu -i /tmp/$1.ncu-rep --page source --csv > gpurun_out/$1_source.csv 2>/dev/null
}
# decision v4: launches per step = pilot(K2), policy, K2 main, K4 (skipped), K4 queue; second step -> skip 5+... take the big K=2 launch of step 2
cap r02_k2 tpr_kernel 5 decision v4 10000000
cap r02_k4 tpr_kernel 1 single v4 10000000
cap r02_k6 tpr_kernel 1 single v3v4 10000000
ls -la gpurun_out/r02_k*
