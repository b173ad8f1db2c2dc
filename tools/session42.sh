#!/bin/bash
# Round 2, GPU session 42: register trims of the persistent traversal loop (leaf ref walks its triangles / current-space direction in the cold
# shared-memory state / direction signs from the sign bits of 1/d), singly and together, against the committed kernel
mkdir -p gpurun_out
T0=$(date +%s)
timeout 600 python -m pytest tests/test_gpu_trace_parity.py tests/test_gpu_render_parity.py tests/test_zgpu_random_scenes.py tests/test_emitters.py -q -m gpu -x 2>&1 | tail -2
export AB_WORKLOADS="C1:64 C2:32 C3:16 C4:4"
tools/ab.sh lib_prev.so libbarnacle_b200.so lib_trim_k.so lib_trim_kd.so lib_trim_ks.so lib_trim_ds.so lib_prev.so libbarnacle_b200.so
echo "== done after $(( $(date +%s) - T0 )) s"
