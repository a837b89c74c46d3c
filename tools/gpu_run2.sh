#!/bin/bash
# round-2 development run: tests, quick bench, launch lists of the exact-mode steps
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02_tests_c.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02_tests_c.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --no-cli > gpurun_out/r02_bench_c.json 2> gpurun_out/r02_bench_c.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_c.json'))
r=d['roofline']
print('C2 value %.4g frac %.3f single %.4g (%.3f) slab_only %.4g exact %.4g (%.3f)'%(d['value'],r['frac'],r['single_sweep_value'],r['single_sweep_frac'],r['slab_only_value'],d['exact_ee']['value'],d['roofline_exact']['frac']))
for k,c in d['configs'].items():
    print(k,'dec %.4g (%.3f) exact %.4g (%.3f) esc %.3f'%(c['decision']['value'],c['decision']['frac'],c['exact_ee']['value'],c['exact_ee']['frac'],c['escalated_fraction']))
PY
for w in "10000000 exact v4" "10000000 exact real" "10000000 exact v3v4" "10000000 decision real" "2000000 exact ccs"; do
  tag=$(echo $w | tr ' ' '_')
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_$tag.csv python tools/one_step.py $w > gpurun_out/r02_launches_$tag.log 2>&1
  echo "ncu $tag rc=$?"
done
