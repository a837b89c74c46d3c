#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_cli.py tests/test_gz.py -x -q 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_final.json'))
print('C2 value %.4g frac %.3f exact %.4g e2e %.4g'%(d['value'],d['roofline']['frac'],d['exact_ee']['value'],d['e2e']['value']))
for k,c in d['configs'].items():
    print(k,'dec %.4g exact %.4g'%(c['decision']['value'],c['exact_ee']['value']), 'parity', c['parity'].get('ee_bit_mismatches'), c['parity'].get('decision_mode_decision_mismatches'))
print(json.dumps(d['e2e_cli']['compressed'], indent=1))
print(d['e2e_cli']['collapse_default'], d['e2e_cli']['no_collapse_fastq'])
PY
grep -o "^\[bench *[0-9.]*s\] [A-Za-z0-9_]*" gpurun_out/r02_bench_final.err | tail -2
