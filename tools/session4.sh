#!/bin/bash
# Round 2, GPU session 4: bank-swizzled 4-wide nodes.  Parity, A/B vs binary and vs the 256-bit-load / 8-CTA variants, ncu.
T0=$(date +%s)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3 | tee gpurun_out/r02c_pytest_gpu.txt
echo "== tests done after $(( $(date +%s) - T0 )) s"
AB_WORKLOADS="C1:64 C2:32 C4:4 C3:16" tools/ab.sh libbarnacle_b200.so lib_ldg256.so lib_8cta.so lib_ldg256_8cta.so
echo "== binary nodes"
for w in C1:64 C2:32 C4:4 C3:16; do BN_BINARY_NODES=1 timeout 300 python bench.py --workload ${w%%:*} --spp ${w##*:} --steps 2 --warmup 1 --no-cpu-baseline --no-configs 2>/dev/null | python tools/benchsum.py; done
for v in lib_ldg256.so; do echo "== parity with $v"; BN_LIB=$PWD/barnacle_b200/lib/$v timeout 300 python -m pytest tests/test_gpu_trace_parity.py tests/test_gpu_render_parity.py tests/test_zgpu_random_scenes.py -q -m gpu -x 2>&1 | tail -2; done
echo "== A/B done after $(( $(date +%s) - T0 )) s"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:^k_traverse$ -s 2 -c 4 -o gpurun_out/r02c_traverse -f \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-configs > gpurun_out/r02c_prof_traverse.log 2>&1
timeout 400 ncu --set full --clock-control none -k regex:^k_traverse$ -s 2 -c 2 -o gpurun_out/r02c_traverse_C4 -f \
  python bench.py --workload C4 --spp 8 --steps 1 --warmup 0 --no-cpu-baseline --no-configs > gpurun_out/r02c_prof_traverse_C4.log 2>&1
echo "== done after $(( $(date +%s) - T0 )) s"
