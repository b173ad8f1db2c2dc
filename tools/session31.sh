#!/bin/bash
# Round 2, GPU session 31: k_sort_rank with its key loads batched (parity of the ordering + the frame's "other" class)
mkdir -p gpurun_out
T0=$(date +%s)
timeout 900 python -m pytest tests/test_gpu_render_parity.py tests/test_gpu_fullsize.py tests/test_emitters.py -q -m gpu -x 2>&1 | tail -3
export AB_WORKLOADS="C1:64 C2:32 C3:16 C4:4"
tools/ab.sh libbarnacle_b200.so libbarnacle_b200.so
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_sort_ -c 12 --csv python bench.py --spp 32 --steps 1 --warmup 0 --no-cpu-baseline --no-configs 2>/dev/null | grep k_sort | cut -d, -f5,15- | head -12
echo "== done after $(( $(date +%s) - T0 )) s"
