#!/bin/bash
for v in default plaincnt default plaincnt; do
echo "== $v"
if [ $v = default ]; then unset MOIRA_B200_LIB; else export MOIRA_B200_LIB=$PWD/build/variants/lib_$v.so; fi
timeout 300 python tools/exact_ab.py v4:10000000 real:10000000 v3v4:10000000 2>&1 | sed 's/first pass [0-9.]* ms//g'
done 2>&1 | tee gpurun_out/r02_cnt_ab.txt
