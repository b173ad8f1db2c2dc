#!/bin/bash
# First GPU call of the next session: A/B of the four default-off experiments against the
# committed build, then the hit / render parity tests on each variant.  Build the variants on the CPU box first:
#   python -m barnacle_b200.build --define=BN_EXP_ANY_UNORDERED --out=lib_anyun.so
#   python -m barnacle_b200.build --define=BN_EXP_SHARED_RCP --out=lib_srcp.so
#   python -m barnacle_b200.build --define=BN_EXP_STAY_REFILL=14 --out=lib_stayrf.so
#   python -m barnacle_b200.build --define=BN_EXP_SCAN_LEAF --out=lib_scanleaf.so
#   python -m barnacle_b200.build --define=BN_EXP_SHARED_RCP --define=BN_EXP_STAY_REFILL=14 --define=BN_EXP_SCAN_LEAF --out=lib_both.so
#   python -m barnacle_b200.build --force          # the committed library last, so that it is the newest
# usage (on the GPU box): tools/next_session.sh > gpurun_out/next_session.log 2>&1
T0=$(date +%s)
AB_WORKLOADS="C2:32 C4:4 C3:16 C1:64" tools/ab.sh libbarnacle_b200.so lib_stayrf.so lib_scanleaf.so lib_anyun.so lib_srcp.so lib_both.so
echo "== A/B done after $(( $(date +%s) - T0 )) s"
for v in lib_stayrf.so lib_scanleaf.so lib_anyun.so lib_srcp.so; do
  [ -f barnacle_b200/lib/$v ] || continue
  echo "== parity with $v"
  BN_LIB=$PWD/barnacle_b200/lib/$v timeout 200 python -m pytest tests -q -m gpu -x 2>&1 | tail -2
done
echo "== done after $(( $(date +%s) - T0 )) s"
