#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k tpr_kernel -s 7 -c 1 -o /tmp/r02_cls python tools/one_step.py 10000000 exact v4 > gpurun_out/r02_cls.log 2>&1
echo "ncu rc=$?"
ncu -i /tmp/r02_cls.ncu-rep --page raw --csv > gpurun_out/r02_cls_raw.csv 2>/dev/null
ncu -i /tmp/r02_cls.ncu-rep --page source --csv > gpurun_out/r02_cls_source.csv 2>/dev/null
