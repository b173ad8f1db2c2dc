#!/bin/bash
# Round 2, GPU session 35: phase-loop thresholds once more on the final kernels (cell-major key, swizzled nodes)
mkdir -p gpurun_out
T0=$(date +%s)
export AB_WORKLOADS="C1:64 C2:32 C3:16 C4:4"
tools/ab.sh libbarnacle_b200.so lib_r16.so lib_r24.so lib_s8.so lib_t6.so libbarnacle_b200.so
echo "== done after $(( $(date +%s) - T0 )) s"
