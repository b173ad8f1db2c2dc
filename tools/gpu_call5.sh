#!/bin/bash
mkdir -p gpurun_out
{
echo "== pytest gpu (all)"
timeout 700 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
echo "== frames A/B (second line of each pair: BN_NO_FRAMES=1)"
for w in C2:32 C1:64 C3:16 C4:4; do
  timeout 300 python bench.py --workload ${w%%:*} --spp ${w##*:} --steps 2 --warmup 1 --no-cpu-baseline 2>/dev/null | python tools/benchsum.py
  BN_NO_FRAMES=1 timeout 300 python bench.py --workload ${w%%:*} --spp ${w##*:} --steps 2 --warmup 1 --no-cpu-baseline 2>/dev/null | python tools/benchsum.py
done
timeout 100 python tools/mlt_bench.py 2>&1 | tail -2
BN_NO_FRAMES=1 timeout 100 python tools/mlt_bench.py 2>&1 | tail -2
} > gpurun_out/call5.log 2>&1
tail -40 gpurun_out/call5.log
