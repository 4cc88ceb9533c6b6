set -x
timeout 120 python profiles/parity_quick.py 2>&1 | tail -1
timeout 200 python profiles/chain_bench.py 2>&1 | tail -3
timeout 200 python profiles/chain_bench.py --players 9 --steps 100 2>&1 | tail -3
timeout 200 python profiles/chain_bench.py --players 4 --steps 100 2>&1 | tail -3
PKB_TIMING_CHAIN=8 PKB_LIB=$PWD/pokerl_b200/_lib/variants/timing.so timeout 200 python profiles/warp_timing.py 2>&1 | tail -10
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
