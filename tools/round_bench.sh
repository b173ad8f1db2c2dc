#!/bin/bash
# Round-end evidence run (one GPU): parity tests, every BASELINE config, the reference arm, the ncu
# launch list and one --set full capture of the traversal and shade kernels.  Outputs in gpurun_out/.
# usage: tools/round_bench.sh <round tag, e.g. r01>
TAG=${1:-r01}
mkdir -p gpurun_out
timeout 500 python -m pytest tests -q -m gpu -x 2>&1 | tail -1 | tee gpurun_out/${TAG}_pytest_gpu.txt
timeout 300 python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_C2_n1.json 2> gpurun_out/${TAG}_bench_C2_n1.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_C2_reference.json 2>/dev/null
timeout 200 python bench.py --workload C1 --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_C1_n1.json 2>/dev/null
timeout 400 python bench.py --workload C3 --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_C3_n1.json 2>/dev/null
timeout 600 python bench.py --workload C4 --steps 1 --warmup 1 > gpurun_out/${TAG}_bench_C4_n1.json 2>/dev/null
timeout 200 python bench.py --workload C5 --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_C5_n1.json 2>/dev/null
for c in 2 1 3 4; do python tools/benchsum.py < gpurun_out/${TAG}_bench_C${c}_n1.json; done
# rays per launch of the full-size frame (for the per-ray DRAM traffic of the profiled launches)
BN_DEBUG_COUNTS=1 timeout 200 python bench.py --steps 1 --warmup 0 --no-cpu-baseline 2>&1 >/dev/null | grep bn_counts | head -4 > gpurun_out/${TAG}_counts.txt
# launch list: cold-cache, serialised -> compare SHARES with the event-timed classes of the bench line
timeout 280 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --spp 8 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_launch.log 2>&1
# --set full of the traversal kernel at the FULL config: extend and shadow launches of bounce 1 of the first 64 Mi-path wave
timeout 400 ncu --set full --clock-control none --import-source on -k regex:^k_traverse$ -s 2 -c 2 -o gpurun_out/${TAG}_traverse -f \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/${TAG}_prof_traverse.log 2>&1
timeout 280 ncu --set full --clock-control none --import-source on -k regex:^k_shade$ -s 1 -c 1 -o gpurun_out/${TAG}_shade -f \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/${TAG}_prof_shade.log 2>&1
ls -la gpurun_out | grep ${TAG}_ | awk '{print $5, $9}'
