#!/bin/bash
# tile-shape sweep for the larger tables: every variant at P=9 (side-pot-heavy mix) and P=4
for lib in pokerl_b200/_lib/variants/*.so; do
  export PKB_LIB=$PWD/$lib
  timeout 120 python profiles/parity_quick.py 2>&1 | tail -1
  for P in 9 4; do
    PKB_VERBOSE=1 timeout 120 python profiles/quick_step_bench.py --steps 2 --players $P 2>&1 | grep "pkb:" | head -1
    timeout 120 python profiles/quick_step_bench.py --steps 60 --players $P 2>&1 | tail -1
    timeout 120 python profiles/overlap_bench.py --steps 300 --players $P 2>&1 | tail -1
  done
done
