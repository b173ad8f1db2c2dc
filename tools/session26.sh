#!/bin/bash
# Round 2, GPU session 26: the ordered (now latency-bound) traversal kernel at 10 / 8 CTAs per SM (48 / 64 registers) and with swizzled node chunks
mkdir -p gpurun_out
T0=$(date +%s)
export AB_WORKLOADS="C1:64 C2:32 C3:16 C4:4"
tools/ab.sh libbarnacle_b200.so lib_tb10.so lib_tb8.so lib_swz.so libbarnacle_b200.so
echo "== done after $(( $(date +%s) - T0 )) s"
