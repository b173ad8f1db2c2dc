#!/bin/bash
# Round 2, GPU session 16: parity + A/B of the SEGMENTED path ordering (state moved by its own kernel inside 256 Ki-path segments) against the
# move fused into extend's refill and against no ordering
mkdir -p gpurun_out
T0=$(date +%s)
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
echo "== tests done after $(( $(date +%s) - T0 )) s"
export AB_WORKLOADS="C1:64 C2:32 C3:16 C4:4"
echo "== BN_SORT=0"; BN_SORT=0 tools/ab.sh libbarnacle_b200.so
echo "== segmented move, scatter grid = 2 x SMs"; BN_SORT=1 BN_SORT_GRID_MULT=2 tools/ab.sh libbarnacle_b200.so
echo "== segmented move, scatter grid = 4 x SMs"; BN_SORT=1 BN_SORT_GRID_MULT=4 tools/ab.sh libbarnacle_b200.so
echo "== segmented move, scatter grid = 1 x SMs"; BN_SORT=1 BN_SORT_GRID_MULT=1 tools/ab.sh libbarnacle_b200.so
echo "== move inside extend"; BN_SORT=2 tools/ab.sh libbarnacle_b200.so
echo "== A/B done after $(( $(date +%s) - T0 )) s"
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-configs 2>/dev/null | python tools/benchsum.py
echo "== done after $(( $(date +%s) - T0 )) s"
