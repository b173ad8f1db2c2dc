#!/bin/bash
# Round 2, GPU session 28: 64-B triangle records fetched with two 256-bit loads in the closest-hit kernel (default) against the 48-B records
mkdir -p gpurun_out
T0=$(date +%s)
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
echo "== tests done after $(( $(date +%s) - T0 )) s"
export AB_WORKLOADS="C1:64 C2:32 C3:16 C4:4"
tools/ab.sh lib_tri48.so libbarnacle_b200.so lib_tri48.so libbarnacle_b200.so
echo "== done after $(( $(date +%s) - T0 )) s"
