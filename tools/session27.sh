#!/bin/bash
# Round 2, GPU session 27: swizzled node chunks as the default (parity suite), against the plain layout; split-phase refill on top
mkdir -p gpurun_out
T0=$(date +%s)
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
echo "== tests done after $(( $(date +%s) - T0 )) s"
export AB_WORKLOADS="C1:64 C2:32 C3:16 C4:4"
tools/ab.sh lib_noswz.so libbarnacle_b200.so lib_split.so lib_noswz.so libbarnacle_b200.so lib_split.so
echo "== done after $(( $(date +%s) - T0 )) s"
