#!/bin/bash
# Round 2, GPU session 43: five cold words per thread instead of nine (the traversal kernels fit the 32-KB shared-memory carve-out: L1 224 KB
# instead of 192 KB) against the committed kernel; full parity suite first
mkdir -p gpurun_out
T0=$(date +%s)
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -2
echo "== tests done after $(( $(date +%s) - T0 )) s"
export AB_WORKLOADS="C1:64 C2:32 C3:16 C4:4"
tools/ab.sh lib_prev.so libbarnacle_b200.so lib_prev.so libbarnacle_b200.so
echo "== done after $(( $(date +%s) - T0 )) s"
