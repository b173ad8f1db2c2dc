#!/bin/bash
# parity on the GPU (whole suite) + every BASELINE config at full size, no profiler
TAG=r01b
mkdir -p gpurun_out
timeout 700 python -m pytest tests -q -m gpu -x 2>&1 | tail -3 | tee gpurun_out/${TAG}_pytest_gpu.txt
timeout 300 python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_C2_n1.json 2> gpurun_out/${TAG}_bench_C2_n1.err
timeout 200 python bench.py --workload C1 --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_C1_n1.json 2>/dev/null
timeout 400 python bench.py --workload C3 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_bench_C3_n1.json 2>/dev/null
timeout 600 python bench.py --workload C4 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_bench_C4_n1.json 2>/dev/null
timeout 200 python bench.py --workload C5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_C5_n1.json 2>/dev/null
for c in 2 1 3 4; do python tools/benchsum.py < gpurun_out/${TAG}_bench_C${c}_n1.json; done
tail -c 600 gpurun_out/${TAG}_bench_C5_n1.json
