#!/bin/bash
for v in split8 split11 split14 split8 split11 split14; do
echo "== $v"; MOIRA_B200_LIB=$PWD/build/variants/lib_$v.so timeout 300 python tools/exact_ab.py v4:10000000 ccs:2000000 v3v4:10000000 2>&1 | sed 's/decision.*| exact/exact/'
done
