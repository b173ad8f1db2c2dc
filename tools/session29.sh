#!/bin/bash
# Round 2, GPU session 29: compute-sanitizer on every device entry point with the ordering kernels in the path, then the round's final evidence run
T0=$(date +%s)
mkdir -p gpurun_out
for tool in memcheck racecheck initcheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize.py > gpurun_out/r02g_sanitizer_$tool.log 2>&1
  echo "== compute-sanitizer $tool: rc=$? after $(( $(date +%s) - T0 )) s"; tail -3 gpurun_out/r02g_sanitizer_$tool.log
done
tools/round_refresh.sh r02g
echo "== all done after $(( $(date +%s) - T0 )) s"
