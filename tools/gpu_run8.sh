#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02_tests_h.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02_tests_h.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_d.json 2> gpurun_out/r02_bench_d.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_d.json'))
r=d['roofline']
print('C2 value %.4g frac %.3f single %.4g (%.3f) slab_only %.4g exact %.4g (%.3f) e2e %.4g'%(d['value'],r['frac'],r['single_sweep_value'],r['single_sweep_frac'],r['slab_only_value'],d['exact_ee']['value'],d['roofline_exact']['frac'],d['e2e']['value']))
for k,c in d['configs'].items():
    print(k,'dec %.4g (%.3f) exact %.4g (%.3f) esc %.3f'%(c['decision']['value'],c['decision']['frac'],c['exact_ee']['value'],c['exact_ee']['frac'],c['escalated_fraction']))
print(json.dumps(d['configs']['C4']['collapse'])[:1500])
print(json.dumps(d['e2e_cli'])[:1200])
PY


