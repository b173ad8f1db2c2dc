#!/bin/bash
# Round 2, GPU session 1: parity of the new tests + folded fix-up, the new bench line, A/B of the queued experiments,
# ncu captures incl. the deep bounces.  usage (GPU box): tools/session1.sh > gpurun_out/s1.log 2>&1
T0=$(date +%s)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader; nproc
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -4 | tee gpurun_out/r02a_pytest_gpu.txt
echo "== tests done after $(( $(date +%s) - T0 )) s"
timeout 400 python bench.py --steps 3 --warmup 3 > gpurun_out/r02a_bench_C2_n1.json 2> gpurun_out/r02a_bench_C2_n1.err; tail -3 gpurun_out/r02a_bench_C2_n1.err
python tools/benchsum.py < gpurun_out/r02a_bench_C2_n1.json
echo "== bench done after $(( $(date +%s) - T0 )) s"
AB_WORKLOADS="C1:64 C2:32 C4:4 C3:16" tools/ab.sh lib_nodrain.so libbarnacle_b200.so lib_stayrf.so lib_scanleaf.so lib_anyun.so lib_srcp.so
echo "== separate fix-up launches with the committed library"
for w in C1:64 C2:32; do BN_SEPARATE_FIXUP=1 timeout 300 python bench.py --workload ${w%%:*} --spp ${w##*:} --steps 2 --warmup 1 --no-cpu-baseline --no-configs 2>/dev/null | python tools/benchsum.py; done
echo "== A/B done after $(( $(date +%s) - T0 )) s"
BN_DEBUG_COUNTS=1 timeout 200 python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-configs 2>&1 >/dev/null | grep bn_counts | head -4 > gpurun_out/r02a_counts.txt
timeout 280 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02a_launches.csv \
  python bench.py --spp 8 --steps 1 --warmup 1 --no-cpu-baseline --no-configs > gpurun_out/r02a_launch.log 2>&1
# extend + shadow launches of bounces 1..4 of the first wave of the full-size frame
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_traverse$ -s 2 -c 8 -o gpurun_out/r02a_traverse -f \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-configs > gpurun_out/r02a_prof_traverse.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:^k_shade$ -s 1 -c 1 -o gpurun_out/r02a_shade -f \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-configs > gpurun_out/r02a_prof_shade.log 2>&1
ls -la gpurun_out | awk '{print $5, $9}'
echo "== done after $(( $(date +%s) - T0 )) s"
