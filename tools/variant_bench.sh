#!/bin/bash
# tuning helper: run bench.py (kernel-only) against each experimental library under build/variants/
for f in build/variants/lib_*.so; do
  echo "== $f"
  MOIRA_B200_LIB=$PWD/$f timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.4g  frac %.3f  kernel_ms %.4f  exact %.4g' % (d['value'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['exact_ee']['value']))"
done
