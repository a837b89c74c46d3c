#!/bin/bash
mkdir -p gpurun_out
for v in uni0 uni8 uni0 uni8; do
echo "== $v"; MOIRA_B200_LIB=$PWD/build/variants/lib_$v.so timeout 300 python tools/exact_ab.py v4:10000000 real:10000000 2>&1 | tee -a gpurun_out/r02_exact_ab_$v.txt
done
MOIRA_B200_LIB=$PWD/build/variants/lib_uni8.so timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
