#!/bin/bash
# Round 2, GPU session 21: the permutation's gathers past L1 (ld.global.cg, default) against cached (lib_ca.so); streaming stores of the ordered planes (lib_stcs.so)
mkdir -p gpurun_out
T0=$(date +%s)
timeout 600 python -m pytest tests/test_gpu_render_parity.py tests/test_gpu_fullsize.py -q -m gpu -x 2>&1 | tail -2
export AB_WORKLOADS="C1:64 C2:32 C3:16 C4:4"
tools/ab.sh lib_ca.so libbarnacle_b200.so lib_stcs.so lib_ca.so libbarnacle_b200.so
echo "== done after $(( $(date +%s) - T0 )) s"
