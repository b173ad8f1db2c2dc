#!/bin/bash
# Round 2, GPU session 17: the ordering's variants (third plane moved by extend | gathered by shade; first ordered bounce 1 | 2) and the phase-loop
# thresholds re-swept on ordered rays
mkdir -p gpurun_out
T0=$(date +%s)
timeout 900 python -m pytest tests/test_gpu_render_parity.py tests/test_gpu_fullsize.py tests/test_emitters.py -q -m gpu -x 2>&1 | tail -3
echo "== tests done after $(( $(date +%s) - T0 )) s"
export AB_WORKLOADS="C1:64 C2:32 C3:16 C4:4"
echo "== BN_SORT=0"; BN_SORT=0 tools/ab.sh libbarnacle_b200.so
echo "== mode 2 (extend moves all three planes)"; BN_SORT=2 tools/ab.sh libbarnacle_b200.so
echo "== mode 3 (extend moves two planes, shade gathers the third)"; BN_SORT=3 tools/ab.sh libbarnacle_b200.so
echo "== mode 2 from bounce 2"; BN_SORT=2 BN_SORT_FROM=2 tools/ab.sh libbarnacle_b200.so
export AB_WORKLOADS="C2:32 C3:16 C4:4"
tools/ab.sh lib_r10.so lib_r18.so lib_s4.so lib_s10.so lib_t2.so lib_t8.so libbarnacle_b200.so
echo "== done after $(( $(date +%s) - T0 )) s"
