#!/bin/bash
# Short evidence refresh after a traversal-only change (one GPU, ~4 min): parity tests, C2, the --set full
# capture of the traversal kernel at full size, the launch list, rays per launch, then C4 / C3 / C1.
# The reference arm, C5 and the shade capture are untouched by such a change: tools/round_bench.sh runs everything.
# usage: tools/round_refresh.sh <tag, e.g. r01e>
TAG=${1:-r01e}
T0=$(date +%s)
stamp() { echo "== $1 done after $(( $(date +%s) - T0 )) s" >&2; }
mkdir -p gpurun_out
timeout 120 python -m pytest tests -q -m gpu -x 2>&1 | tail -1 | tee gpurun_out/${TAG}_pytest_gpu.txt; stamp pytest
timeout 60 python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_C2_n1.json 2> gpurun_out/${TAG}_bench_C2_n1.err; stamp C2
python tools/benchsum.py < gpurun_out/${TAG}_bench_C2_n1.json
timeout 90 ncu --set full --clock-control none --import-source on -k regex:^k_traverse$ -s 2 -c 2 -o gpurun_out/${TAG}_traverse -f \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/${TAG}_prof_traverse.log 2>&1; stamp ncu_traverse
timeout 60 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --spp 8 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_launch.log 2>&1; stamp launches
BN_DEBUG_COUNTS=1 timeout 40 python bench.py --steps 1 --warmup 0 --no-cpu-baseline 2>&1 >/dev/null | grep bn_counts | head -4 > gpurun_out/${TAG}_counts.txt; stamp counts
timeout 100 python bench.py --workload C4 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_bench_C4_n1.json 2>/dev/null; stamp C4
timeout 60 python bench.py --workload C3 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_bench_C3_n1.json 2>/dev/null; stamp C3
timeout 40 python bench.py --workload C1 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_C1_n1.json 2>/dev/null; stamp C1
for c in 4 3 1; do python tools/benchsum.py < gpurun_out/${TAG}_bench_C${c}_n1.json; done
