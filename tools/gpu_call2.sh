#!/bin/bash
mkdir -p gpurun_out
{
echo "== traversal A/B 2"
AB_WORKLOADS="C2:32 C4:4 C3:16" tools/ab.sh libbarnacle_b200.so lib_T4.so lib_T2.so lib_T1.so lib_T4N8.so lib_T4N16.so lib_T4RA6.so lib_T4RA18.so
} > gpurun_out/call2.log 2>&1
tail -80 gpurun_out/call2.log
