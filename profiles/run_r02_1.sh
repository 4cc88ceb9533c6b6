set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 120 python profiles/parity_quick.py 2>&1 | tail -2
timeout 200 python profiles/chain_bench.py 2>&1 | tail -3
PKB_NO_PDL=1 timeout 200 python profiles/chain_bench.py 2>&1 | tail -3
for v in w2 w1; do
  export PKB_LIB=$PWD/pokerl_b200/_lib/variants/$v.so
  timeout 120 python profiles/parity_quick.py 2>&1 | tail -1
  timeout 200 python profiles/chain_bench.py 2>&1 | tail -3
done
PKB_LIB=$PWD/pokerl_b200/_lib/variants/timing.so timeout 200 python profiles/warp_timing.py 2>&1 | tail -16
unset PKB_LIB
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
