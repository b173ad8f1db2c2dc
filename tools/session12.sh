#!/bin/bash
# Round 2, GPU session 12: (a) parity + A/B of the candidate pre-pass kernel; (b) what sorting the live paths per bounce is worth (experiment build, library sort)
mkdir -p gpurun_out
T0=$(date +%s)
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
echo "== tests done after $(( $(date +%s) - T0 )) s"
AB_WORKLOADS="C1:64 C2:32 C3:16 C4:4" tools/ab.sh lib_nocand.so libbarnacle_b200.so lib_nocand.so libbarnacle_b200.so
echo "== A/B done after $(( $(date +%s) - T0 )) s"
for mb in 0 2 3 4 5; do
  echo "== sorted per bounce, octant + morton bits/axis = $mb"
  BN_SORT=1 BN_SORT_MBITS=$mb AB_WORKLOADS="C2:32 C3:16 C4:4" tools/ab.sh lib_expsort.so
done
echo "== unsorted, same build"; AB_WORKLOADS="C2:32 C3:16 C4:4" tools/ab.sh lib_expsort.so
echo "== sort from bounce 2 only (mbits 3)"; BN_SORT=1 BN_SORT_FROM=2 AB_WORKLOADS="C2:32" tools/ab.sh lib_expsort.so
echo "== done after $(( $(date +%s) - T0 )) s"
