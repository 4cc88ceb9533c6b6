timeout 120 python profiles/parity_quick.py 2>&1 | tail -1
timeout 200 python profiles/chain_bench.py 2>&1 | tail -3
export PKB_LIB=$PWD/pokerl_b200/_lib/variants/relaxed.so
timeout 120 python profiles/parity_quick.py 2>&1 | tail -1
timeout 200 python profiles/chain_bench.py 2>&1 | tail -3
unset PKB_LIB
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
