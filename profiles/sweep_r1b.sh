#!/bin/bash
# A/B of the step kernel forms on one box: block-tile (PKB_STEP_IMPL=tile) vs warp-tile variants.
echo "== block-tile (v9)"; PKB_STEP_IMPL=tile timeout 120 python profiles/parity_quick.py 2>&1 | tail -1
PKB_STEP_IMPL=tile timeout 120 python profiles/quick_step_bench.py --steps 100 2>&1 | tail -1
PKB_STEP_IMPL=tile timeout 120 python profiles/quick_step_bench.py --steps 50 --players 9 2>&1 | tail -1
echo "== warp-tile variants"
PKB_VERBOSE=1 timeout 120 python profiles/quick_step_bench.py --steps 2 2>&1 | grep "pkb:" | head -1
SWEEP_P9=1 bash profiles/sweep.sh
