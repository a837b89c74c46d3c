#!/bin/bash
# round-2 final evidence, one GPU: reference arm, bench, launch list of the bench command, ncu captures of the classify-first path
mkdir -p gpurun_out
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; echo "reference rc=$?"
timeout 900 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_final.json'))
r=d['roofline']
print('C2 value %.4g frac %.3f single %.4g (%.3f) exact %.4g (%.3f) e2e %.4g launches %d'%(d['value'],r['frac'],r['single_sweep_value'],r['single_sweep_frac'],d['exact_ee']['value'],d['roofline_exact']['frac'],d['e2e']['value'],d['gpu_launches']))
for k,c in d['configs'].items():
    print(k,'dec %.4g (%.3f) exact %.4g (%.3f) esc %.3f cf %s'%(c['decision']['value'],c['decision']['frac'],c['exact_ee']['value'],c['exact_ee']['frac'],c['escalated_fraction'],c.get('classified_first_fraction')), 'parity', c['parity'].get('ee_bit_mismatches'), c['parity'].get('decision_mismatches_outside_band'), c['parity'].get('decision_mode_lower_bound_violations'))
print(json.dumps(d['configs']['C4']['collapse'])[:600])
print(json.dumps(d['e2e_cli'])[:2500])
print(d['clocks'])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tpr_kernel|ladder|wpr_|blk_|policy|count_marks|sorted_first|len_|grp_|dedup|seq_hash|fp64_peak|unpack|Device|fq_|contig" -c 3000 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-cli > gpurun_out/r02_launches.log 2>&1; echo "launch list rc=$?"
cap() {  # tag kernel-regex skip count mode profile n source?
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -o /tmp/$1 python tools/one_step.py $7 $5 $6 > gpurun_out/$1.log 2>&1
  echo "ncu $1 rc=$?"
  ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
  if [ "$8" = "src" ]; then ncu -i /tmp/$1.ncu-rep --page source --csv > gpurun_out/$1_source.csv 2>/dev/null; fi
}
cap r02_cf_classifier tpr_kernel 2 1 decision ccs 2000000 src
cap r02_cf_ladder_dec ladder_tpr 2 2 decision ccs 2000000
cap r02_cf_ladder_exact ladder_tpr 2 2 exact ccs 2000000
ls -la gpurun_out/r02_cf_*
