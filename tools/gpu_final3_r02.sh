#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r02_tests_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_final.json'))
r=d['roofline']
print('C2 value %.4g frac %.3f single %.4g (%.3f) exact %.4g (%.3f) e2e %.4g launches %d'%(d['value'],r['frac'],r['single_sweep_value'],r['single_sweep_frac'],d['exact_ee']['value'],d['roofline_exact']['frac'],d['e2e']['value'],d['gpu_launches']))
for k,c in d['configs'].items():
    print(k,'dec %.4g (%.3f) exact %.4g (%.3f) esc %.3f cf %s'%(c['decision']['value'],c['decision']['frac'],c['exact_ee']['value'],c['exact_ee']['frac'],c['escalated_fraction'],c.get('classified_first_fraction')), 'parity', c['parity'].get('ee_bit_mismatches'), c['parity'].get('decision_mismatches_outside_band'), c['parity'].get('decision_mode_decision_mismatches'), c['parity'].get('decision_mode_lower_bound_violations'))
print(json.dumps(d['e2e_cli'])[:1800])
print(d['clocks'])
PY
grep -o "^\[bench *[0-9.]*s\] [A-Za-z0-9_]*" gpurun_out/r02_bench_final.err | tail -3
