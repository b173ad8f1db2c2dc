#!/bin/bash
mkdir -p gpurun_out
{
echo "== traversal A/B 4 (fine sweep around T4 / N6 / R12 / RA16)"
AB_WORKLOADS="C2:32 C4:4 C3:16" tools/ab.sh libbarnacle_b200.so lib_T3.so lib_T5.so lib_N5.so lib_N7.so lib_R10.so lib_R14.so lib_RA14.so lib_RA18.so
} > gpurun_out/call6.log 2>&1
tail -60 gpurun_out/call6.log
