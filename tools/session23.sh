#!/bin/bash
# Round 2, GPU session 23: the claim pipeline over CONSECUTIVE chunks (128 slots; 64 / 256; without the ordering-entry prefetch) against the committed kernel
mkdir -p gpurun_out
T0=$(date +%s)
timeout 900 python -m pytest tests/test_gpu_render_parity.py tests/test_gpu_trace_parity.py tests/test_gpu_fullsize.py tests/test_pssmlt.py -q -m gpu -x 2>&1 | tail -3
echo "== tests done after $(( $(date +%s) - T0 )) s"
export AB_WORKLOADS="C1:64 C2:32 C3:16 C4:4"
tools/ab.sh lib_noclaim.so libbarnacle_b200.so lib_claim_noord.so lib_c64.so lib_c256.so lib_noclaim.so libbarnacle_b200.so
echo "== done after $(( $(date +%s) - T0 )) s"
