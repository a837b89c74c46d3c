#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; echo "reference rc=$?"; cut -c1-400 gpurun_out/r02_bench_reference.json
timeout 900 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_final.json'))
r=d['roofline']
print('C2 value %.4g frac %.3f single %.4g (%.3f) slab_only %.4g exact %.4g (%.3f) e2e %.4g launches %d'%(d['value'],r['frac'],r['single_sweep_value'],r['single_sweep_frac'],r['slab_only_value'],d['exact_ee']['value'],d['roofline_exact']['frac'],d['e2e']['value'],d['gpu_launches']))
for k,c in d['configs'].items():
    print(k,'dec %.4g (%.3f) exact %.4g (%.3f) esc %.3f'%(c['decision']['value'],c['decision']['frac'],c['exact_ee']['value'],c['exact_ee']['frac'],c['escalated_fraction']), 'parity', c['parity'].get('ee_bit_mismatches'), c['parity'].get('decision_mismatches_outside_band'))
print(json.dumps(d['configs']['C4']['collapse'])[:700])
print(json.dumps(d['e2e_cli'])[:1500])
print(d['clocks'])
PY
