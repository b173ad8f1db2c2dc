#!/bin/bash
# Round 2, GPU session 24: ordering key cell-major instead of octant-major; shade at 6 / 8 CTAs per SM on ordered paths
mkdir -p gpurun_out
T0=$(date +%s)
export AB_WORKLOADS="C1:64 C2:32 C3:16 C4:4"
tools/ab.sh libbarnacle_b200.so lib_cellmajor.so lib_sh6.so lib_sh8.so libbarnacle_b200.so
echo "== done after $(( $(date +%s) - T0 )) s"
