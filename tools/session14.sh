#!/bin/bash
# Round 2, GPU session 14: parity + A/B of the INDIRECT path ordering (perm only; extend / shade gather their state): off / 512 bins / 4096 bins
mkdir -p gpurun_out
T0=$(date +%s)
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
echo "== tests done after $(( $(date +%s) - T0 )) s"
export AB_WORKLOADS="C1:64 C2:32 C3:16 C4:4"
echo "== BN_SORT=0"; BN_SORT=0 tools/ab.sh libbarnacle_b200.so
echo "== ordered, 512 bins"; tools/ab.sh libbarnacle_b200.so
echo "== ordered, 4096 bins"; tools/ab.sh lib_sort3.so
echo "== BN_SORT=0"; BN_SORT=0 tools/ab.sh libbarnacle_b200.so
echo "== ordered, 512 bins"; tools/ab.sh libbarnacle_b200.so
echo "== A/B done after $(( $(date +%s) - T0 )) s"
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-configs 2>/dev/null | python tools/benchsum.py
echo "== done after $(( $(date +%s) - T0 )) s"
