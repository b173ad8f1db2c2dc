#!/bin/bash
# Round 2, GPU session 30: k_shade with a three-chunk claim pipeline and one 64-bit atomic for both appends, against the committed kernel
mkdir -p gpurun_out
T0=$(date +%s)
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
echo "== tests done after $(( $(date +%s) - T0 )) s"
export AB_WORKLOADS="C1:64 C2:32 C3:16 C4:4"
tools/ab.sh lib_shade_prev.so libbarnacle_b200.so lib_shade_prev.so libbarnacle_b200.so
echo "== done after $(( $(date +%s) - T0 )) s"
