#!/bin/bash
# Round-end evidence run on ONE GPU with the committed library: parity tests, the default bench line (C2 + secondary configs),
# every BASELINE config at full size, the reference arm, launch list, ncu --set full captures of the traversal kernels
# (bounces 1-2 of C2 at full size; bounce 1 of C4) and of k_shade, rays per launch.  Outputs in gpurun_out/<tag>_*.
# usage: tools/round_refresh.sh <tag, e.g. r02>
TAG=${1:-r02}
T0=$(date +%s)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader; nproc
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3 | tee gpurun_out/${TAG}_pytest_gpu.txt
echo "== tests done after $(( $(date +%s) - T0 )) s"
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_C2_n1.json 2> gpurun_out/${TAG}_bench_C2_n1.err; tail -2 gpurun_out/${TAG}_bench_C2_n1.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_C2_reference.json 2>/dev/null
timeout 300 python bench.py --workload C1 --steps 5 --warmup 3 --no-configs > gpurun_out/${TAG}_bench_C1_n1.json 2>/dev/null
timeout 500 python bench.py --workload C3 --steps 2 --warmup 1 --no-configs > gpurun_out/${TAG}_bench_C3_n1.json 2>/dev/null
timeout 700 python bench.py --workload C4 --steps 1 --warmup 1 --no-configs > gpurun_out/${TAG}_bench_C4_n1.json 2>/dev/null
timeout 300 python bench.py --workload C5 --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_C5_n1.json 2>/dev/null
for c in 2 1 3 4; do python tools/benchsum.py < gpurun_out/${TAG}_bench_C${c}_n1.json; done
echo "== benches done after $(( $(date +%s) - T0 )) s"
BN_DEBUG_COUNTS=1 timeout 200 python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-configs 2>&1 >/dev/null | grep bn_counts | head -4 > gpurun_out/${TAG}_counts.txt
BN_DEBUG_COUNTS=1 timeout 200 python bench.py --workload C4 --spp 8 --steps 1 --warmup 0 --no-cpu-baseline --no-configs 2>&1 >/dev/null | grep bn_counts | head -2 > gpurun_out/${TAG}_counts_C4.txt
timeout 280 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --spp 8 --steps 1 --warmup 1 --no-cpu-baseline --no-configs > gpurun_out/${TAG}_launch.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:^k_traverse$ -s 2 -c 8 -o gpurun_out/${TAG}_traverse -f \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-configs > gpurun_out/${TAG}_prof_traverse.log 2>&1
timeout 400 ncu --set full --clock-control none -k regex:^k_traverse$ -s 2 -c 2 -o gpurun_out/${TAG}_traverse_C4 -f \
  python bench.py --workload C4 --spp 8 --steps 1 --warmup 0 --no-cpu-baseline --no-configs > gpurun_out/${TAG}_prof_traverse_C4.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:^k_shade$ -s 1 -c 1 -o gpurun_out/${TAG}_shade -f \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-configs > gpurun_out/${TAG}_prof_shade.log 2>&1
timeout 200 ncu --set full --clock-control none -k regex:^k_sort_ -s 3 -c 3 -o gpurun_out/${TAG}_sort -f \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-configs > gpurun_out/${TAG}_prof_sort.log 2>&1
ls -la gpurun_out | grep ${TAG}_ | awk '{print $5, $9}'
echo "== done after $(( $(date +%s) - T0 )) s"
