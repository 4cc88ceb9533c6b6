#!/bin/bash
# Runs the parity check and the device-time probes for every library under pokerl_b200/_lib/variants/
# (a -DPKB_TIMING build, if present, runs the per-warp timeline instead).
for lib in pokerl_b200/_lib/variants/*.so; do
  export PKB_LIB=$PWD/$lib
  if [[ $lib == *timing* ]]; then timeout 200 python profiles/warp_timing.py 2>&1 | tail -14; continue; fi
  timeout 120 python profiles/parity_quick.py 2>&1 | tail -1
  timeout 120 python profiles/quick_step_bench.py --steps 100 ${SWEEP_ARGS} 2>&1 | tail -1
  timeout 120 python profiles/overlap_bench.py ${SWEEP_ARGS} 2>&1 | tail -1
done
