#!/bin/bash
# A/B of experiment builds (barnacle_b200/lib/lib_*.so): short bench lines per variant.
# usage: [AB_WORKLOADS="C2:32 C4:4 C3:16"] tools/ab.sh lib_k1.so lib_k2.so ...
WL=${AB_WORKLOADS:-"C2:32 C4:4 C3:16"}
for v in "$@"; do
  echo "== $v"
  for w in $WL; do
    BN_LIB=$PWD/barnacle_b200/lib/$v timeout 300 python bench.py --workload ${w%%:*} --spp ${w##*:} --steps 2 --warmup 1 --no-cpu-baseline --no-configs 2>/dev/null | python tools/benchsum.py
  done
done
