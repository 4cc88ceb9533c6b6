for b in 6 8 12; do for r in 16 32; do
PKB_BENCH_E2E_BATCHES=$b PKB_BENCH_E2E_ROUNDS=$r python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-q-learning 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('batches $b rounds $r e2e %.3e value %.3e ratio %.3f' % (d['e2e']['value'], d['value'], d['e2e']['value']/d['value']))"
done; done
