#!/bin/bash
mkdir -p gpurun_out
{
echo "== traversal A/B 3"
AB_WORKLOADS="C2:32 C4:4 C3:16" tools/ab.sh lib_T4N6.so lib_T4N4.so lib_T4N2.so lib_T4N8RA16.so lib_T4N8RA24.so lib_T4N8R16.so lib_T4N8R8.so
} > gpurun_out/call3.log 2>&1
tail -80 gpurun_out/call3.log
