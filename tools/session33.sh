#!/bin/bash
# Round 2, GPU session 33: the traversal stack's top entry in registers (pop latency off the critical path) against the committed kernel
mkdir -p gpurun_out
T0=$(date +%s)
BN_LIB=$PWD/barnacle_b200/lib/lib_topreg.so timeout 600 python -m pytest tests/test_gpu_trace_parity.py tests/test_gpu_render_parity.py -q -m gpu -x 2>&1 | tail -2
export AB_WORKLOADS="C1:64 C2:32 C3:16 C4:4"
tools/ab.sh libbarnacle_b200.so lib_topreg.so libbarnacle_b200.so lib_topreg.so
echo "== done after $(( $(date +%s) - T0 )) s"
