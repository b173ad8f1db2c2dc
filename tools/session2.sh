#!/bin/bash
# Round 2, GPU session 2: the 4-wide nodes.  Parity (all GPU tests run with the wide nodes by default; the trace / render
# parity files again with BN_BINARY_NODES=1), then A/B wide vs binary on C1..C4.
T0=$(date +%s)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -4 | tee gpurun_out/r02b_pytest_gpu.txt
BN_BINARY_NODES=1 timeout 600 python -m pytest tests/test_gpu_trace_parity.py tests/test_gpu_render_parity.py tests/test_zgpu_random_scenes.py -q -m gpu -x 2>&1 | tail -3
echo "== tests done after $(( $(date +%s) - T0 )) s"
for mode in wide binary; do
  echo "== $mode"
  for w in C1:64 C2:32 C4:4 C3:16; do
    if [ $mode = binary ]; then export BN_BINARY_NODES=1; else unset BN_BINARY_NODES; fi
    timeout 300 python bench.py --workload ${w%%:*} --spp ${w##*:} --steps 2 --warmup 1 --no-cpu-baseline --no-configs 2>/dev/null | python tools/benchsum.py
  done
done
unset BN_BINARY_NODES
echo "== A/B done after $(( $(date +%s) - T0 )) s"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:^k_traverse$ -s 2 -c 4 -o gpurun_out/r02b_traverse -f \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-configs > gpurun_out/r02b_prof_traverse.log 2>&1
echo "== done after $(( $(date +%s) - T0 )) s"
