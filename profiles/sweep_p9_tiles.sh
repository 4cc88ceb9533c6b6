for lib in pokerl_b200/_lib/variants/l*.so; do
  export PKB_LIB=$PWD/$lib
  timeout 120 python profiles/parity_quick.py 2>&1 | tail -1
  PKB_VERBOSE=1 timeout 200 python profiles/chain_bench.py --players 9 --steps 100 2>&1 | grep -E "one handle|blocks/SM" | sort -u | tail -2
done
