#!/bin/bash
# Round 2, GPU session 22: the refill's claim pipeline (two claimed blocks held, the third's atomic in flight, ordering entries fetched a refill ahead)
# against the committed kernel; streaming stores of the ordered planes; 8 Ki / 16 Ki-path tiles in k_sort_rank
mkdir -p gpurun_out
T0=$(date +%s)
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
echo "== tests done after $(( $(date +%s) - T0 )) s"
export AB_WORKLOADS="C1:64 C2:32 C3:16 C4:4"
tools/ab.sh lib_noclaim.so libbarnacle_b200.so lib_claim_stcs.so lib_t8k.so lib_t16k.so lib_noclaim.so libbarnacle_b200.so
echo "== done after $(( $(date +%s) - T0 )) s"
