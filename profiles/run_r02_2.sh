set -x
timeout 120 python profiles/parity_quick.py 2>&1 | tail -1
timeout 200 python profiles/chain_bench.py 2>&1 | tail -3
for v in spin noprefetch; do
  export PKB_LIB=$PWD/pokerl_b200/_lib/variants/$v.so
  timeout 120 python profiles/parity_quick.py 2>&1 | tail -1
  timeout 200 python profiles/chain_bench.py 2>&1 | tail -3
done
for v in timing timing_spin; do
PKB_TIMING_CHAIN=8 PKB_LIB=$PWD/pokerl_b200/_lib/variants/$v.so timeout 200 python profiles/warp_timing.py 2>&1 | tail -17
done
