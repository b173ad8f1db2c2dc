#!/bin/bash
# Round 2, GPU session 6: pre-sorted near/far planes in the flat TLAS; re-tuning of the phase-loop thresholds for the 4-wide nodes.
T0=$(date +%s)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
echo "== tests done after $(( $(date +%s) - T0 )) s"
AB_WORKLOADS="C1:64 C2:32 C4:4 C3:16" tools/ab.sh libbarnacle_b200.so lib_n4.so lib_n8.so lib_n10.so lib_t2.so lib_t6.so lib_r12.so lib_r16.so lib_ra12.so lib_ra20.so
echo "== done after $(( $(date +%s) - T0 )) s"
