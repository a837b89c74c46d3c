for v in 0 1; do
echo "== NO_TMA=$v"
MOIRA_B200_NO_TMA=$v timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.4g  frac %.3f  kernel_ms %.4f  exact %.4g poisson %.4g ee %.4g' % (d['value'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['exact_ee']['value'], d['modes']['poisson']['value'], d['modes']['expected_error']['value']))"
done
