#!/bin/bash
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 300 python tools/exact_ab.py $W 2>&1 | sed 's/first pass [0-9.]* ms//g'; }
{
W="v3v4:10000000"
run MOIRA_B200_CLASSIFY_FIRST_K=1000 MOIRA_B200_CLASSIFY_FIRST_DEC_K=1000
run MOIRA_B200_CLASSIFY_FIRST_K=6 MOIRA_B200_CLASSIFY_FIRST_DEC_K=6
run AB_UNCERT=0.015 MOIRA_B200_CLASSIFY_FIRST_K=1000 MOIRA_B200_CLASSIFY_FIRST_DEC_K=1000
run AB_UNCERT=0.015 MOIRA_B200_CLASSIFY_FIRST_K=8 MOIRA_B200_CLASSIFY_FIRST_DEC_K=8
W="real:10000000 v4:10000000"
run AB_UNCERT=0.03 MOIRA_B200_CLASSIFY_FIRST_K=1000 MOIRA_B200_CLASSIFY_FIRST_DEC_K=1000
run AB_UNCERT=0.03 MOIRA_B200_CLASSIFY_FIRST_K=8 MOIRA_B200_CLASSIFY_FIRST_DEC_K=8
} > gpurun_out/r02_cf_ab4.txt 2>&1
cat gpurun_out/r02_cf_ab4.txt
timeout 900 ncu --set full --clock-control none --import-source on -k tpr_kernel -s 2 -c 1 -o /tmp/r02_cf_cls python tools/one_step.py 2000000 decision ccs > gpurun_out/r02_cf_cls.log 2>&1
echo "ncu rc=$?"
ncu -i /tmp/r02_cf_cls.ncu-rep --page raw --csv > gpurun_out/r02_cf_cls_raw.csv 2>/dev/null
ncu -i /tmp/r02_cf_cls.ncu-rep --page source --csv > gpurun_out/r02_cf_cls_source.csv 2>/dev/null
ls -la gpurun_out/r02_cf_cls*
