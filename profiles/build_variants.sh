#!/bin/bash
# Builds kernel variants into pokerl_b200/_lib/variants/ for profiles/sweep.sh.
#   profiles/build_variants.sh name1:"-DFOO=1 -DBAR=2" name2:"..."
# Each variant is the product source with extra -D flags; the builds run in parallel.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=$ROOT/pokerl_b200/_lib/variants
mkdir -p "$OUT"
for spec in "$@"; do
  name=${spec%%:*}
  flags=${spec#*:}
  [ "$flags" = "$spec" ] && flags=""
  (
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -shared -Xcompiler -fPIC \
      -cudart static -ccbin /usr/bin/g++ $flags -o "$OUT/$name.so" "$ROOT/pokerl_b200/csrc/pkb_kernels.cu" \
      && echo "built $name ($flags)"
  ) &
done
wait
