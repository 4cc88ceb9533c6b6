for i in 1 2 3; do
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-q-learning 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('with sampler   ', round(d['roofline']['frac'],4), [round(x,3) for x in d['timed_regions_ms']])"
PKB_BENCH_NO_CLOCKS=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-q-learning 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('without sampler', round(d['roofline']['frac'],4), [round(x,3) for x in d['timed_regions_ms']])"
done
