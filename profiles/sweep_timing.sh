#!/bin/bash
for lib in pokerl_b200/_lib/variants/t_*.so; do
  export PKB_LIB=$PWD/$lib
  echo "== $lib"
  timeout 120 python profiles/parity_quick.py 2>&1 | tail -1 | cut -c1-12
  timeout 120 python profiles/phase_timing.py 2>&1 | tail -6
  timeout 120 python profiles/quick_step_bench.py --steps 50 2>&1 | tail -1 | cut -d' ' -f2-9
done
