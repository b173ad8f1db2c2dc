#!/bin/bash
# Round 2, GPU session 41: the world-space direction signs read off 1/d when needed instead of living in a register (any-hit kernel's spills 16 -> 8 bytes)
mkdir -p gpurun_out
T0=$(date +%s)
timeout 600 python -m pytest tests/test_gpu_trace_parity.py tests/test_gpu_render_parity.py tests/test_zgpu_random_scenes.py -q -m gpu -x 2>&1 | tail -2
export AB_WORKLOADS="C1:64 C2:32 C3:16 C4:4"
tools/ab.sh lib_prev.so libbarnacle_b200.so lib_prev.so libbarnacle_b200.so
echo "== done after $(( $(date +%s) - T0 )) s"
