#!/bin/bash
# Round 2, GPU session 18: full parity suite on the adopted ordering (extend moves two planes, shade gathers the third), then refill thresholds
# and bin counts re-swept on ordered rays
mkdir -p gpurun_out
T0=$(date +%s)
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
echo "== tests done after $(( $(date +%s) - T0 )) s"
export AB_WORKLOADS="C1:64 C2:32 C3:16 C4:4"
echo "== BN_SORT=0"; BN_SORT=0 tools/ab.sh libbarnacle_b200.so
tools/ab.sh libbarnacle_b200.so lib_r16.so lib_r18.so lib_r20.so lib_r24.so lib_a12.so lib_a20.so lib_m1.so lib_m3.so libbarnacle_b200.so
echo "== A/B done after $(( $(date +%s) - T0 )) s"
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-configs 2>/dev/null | python tools/benchsum.py
echo "== done after $(( $(date +%s) - T0 )) s"
