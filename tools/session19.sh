#!/bin/bash
# Round 2, GPU session 19: new defaults (refill 20 / 20, 4096 bins) confirmed; shadow refill 24, closest refill 22, prefetch distance 8 Ki / 32 Ki
mkdir -p gpurun_out
T0=$(date +%s)
export AB_WORKLOADS="C1:64 C2:32 C3:16 C4:4"
tools/ab.sh libbarnacle_b200.so lib_a24.so lib_r22.so lib_p8k.so lib_p32k.so libbarnacle_b200.so
echo "== done after $(( $(date +%s) - T0 )) s"
