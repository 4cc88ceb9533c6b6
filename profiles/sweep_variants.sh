#!/bin/bash
# profiles/sweep_variants.sh P name1 name2 ...: parity + chain bench of the variants built by build_variants.sh
P=$1; shift
for name in "$@"; do
  export PKB_LIB=$PWD/pokerl_b200/_lib/variants/$name.so
  timeout 120 python profiles/parity_quick.py 2>&1 | tail -1
  timeout 300 python profiles/chain_bench.py --players $P --steps 100 2>&1 | grep -E "one handle|round-robin" | sed 's/lib=//'
done
