#!/bin/bash
# Round 2, GPU session 38: wave size on the ordered wavefront (C2 and C3 at full size / 64 spp)
mkdir -p gpurun_out
T0=$(date +%s)
for wp in 33554432 67108864 134217728; do
  echo "== BN_WAVE_PATHS=$wp"
  BN_WAVE_PATHS=$wp timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs 2>/dev/null | python tools/benchsum.py
  BN_WAVE_PATHS=$wp timeout 300 python bench.py --workload C3 --spp 64 --steps 2 --warmup 1 --no-cpu-baseline --no-configs 2>/dev/null | python tools/benchsum.py
done
echo "== done after $(( $(date +%s) - T0 )) s"
