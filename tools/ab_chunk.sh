for mb in 8 16 32 64 128 256; do
echo "== CHUNK_MB=$mb"; MOIRA_B200_CHUNK_MB=$mb python bench.py --steps 5 --warmup 3 --no-cpu 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('e2e %.4g  ms %.2f  parse %.4g' % (d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e_parse']['value']))"
done
