#!/bin/bash
# Round 2, GPU session 3: L1 gather microbenchmark (LDG.128 vs LDG.256), multi-device C ABI tests, compute-sanitizer runs.
T0=$(date +%s)
mkdir -p gpurun_out
tools/microbench/l1_gather | tee gpurun_out/r02_l1_gather.txt
echo "== microbench done after $(( $(date +%s) - T0 )) s"
timeout 600 python -m pytest tests/test_multi_cabi.py tests/test_gpu_render_parity.py -q -m gpu -x 2>&1 | tail -4
echo "== tests done after $(( $(date +%s) - T0 )) s"
for tool in memcheck racecheck initcheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize.py > gpurun_out/r02_sanitizer_$tool.log 2>&1
  echo "== compute-sanitizer $tool: rc=$? after $(( $(date +%s) - T0 )) s"; tail -4 gpurun_out/r02_sanitizer_$tool.log
done
echo "== done after $(( $(date +%s) - T0 )) s"
