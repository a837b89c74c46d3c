#!/bin/bash
# bench.py at N GPUs, launched like the driver does: bash tools/gpu_scale_r02.sh N
N=$1
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N > gpurun_out/r02_scale_$N.json 2> gpurun_out/r02_scale_$N.err; echo "bench N=$N rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_scale_$N.json').read().strip().splitlines()[-1])
print('N', d['n_gpus'], 'value %.4g exact %.4g e2e %.4g frac_of_link %s'%(d['value'], d['exact_ee']['value'], d['e2e']['value'], d['e2e'].get('frac_of_link')))
for k,c in d['configs'].items():
    print(k,'dec %.4g exact %.4g'%(c['decision']['value'],c['exact_ee']['value']), 'parity', c['parity'].get('ee_bit_mismatches'))
print(json.dumps(d['parity'].get('counters_vs_rank_flags')))
print(json.dumps(d.get('e2e_cli'))[:800])
print(d['clocks'])
PY
grep -o "^\[bench *[0-9.]*s\] [A-Za-z0-9_]*" gpurun_out/r02_scale_$N.err | tail -8
