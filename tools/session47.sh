#!/bin/bash
# Round 2, GPU session 47 (the round's last GPU minutes): two-entry pops in the closest-hit kernel against the committed kernel
mkdir -p gpurun_out
BN_LIB=$PWD/barnacle_b200/lib/lib_pop2.so timeout 200 python -m pytest tests/test_gpu_trace_parity.py tests/test_gpu_render_parity.py -q -m gpu -x 2>&1 | tail -1
export AB_WORKLOADS="C2:32 C4:4 C3:16"
tools/ab.sh libbarnacle_b200.so lib_pop2.so libbarnacle_b200.so lib_pop2.so
