#!/bin/bash
# A/B of experiment builds (barnacle_b200/lib/lib_*.so): short bench lines per variant.
# usage: tools/ab.sh lib_k1.so lib_k2.so ...
for v in "$@"; do
  echo "== $v"
  for w in "C2 32" "C4 4" "C3 16"; do
    set -- $w
    BN_LIB=$PWD/barnacle_b200/lib/$v timeout 300 python bench.py --workload $1 --spp $2 --steps 2 --warmup 1 --no-cpu-baseline 2>/dev/null | python tools/benchsum.py
  done
done
