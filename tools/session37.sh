#!/bin/bash
# Round 2, GPU session 37: the BASELINE configs at full size on the round's final tree (bench lines only)
mkdir -p gpurun_out
T0=$(date +%s)
TAG=r02h
timeout 300 python bench.py --workload C1 --steps 5 --warmup 3 --no-configs > gpurun_out/${TAG}_bench_C1_n1.json 2>/dev/null
timeout 500 python bench.py --workload C3 --steps 2 --warmup 1 --no-configs > gpurun_out/${TAG}_bench_C3_n1.json 2>/dev/null
timeout 700 python bench.py --workload C4 --steps 1 --warmup 1 --no-configs > gpurun_out/${TAG}_bench_C4_n1.json 2>/dev/null
timeout 300 python bench.py --workload C5 --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_C5_n1.json 2>/dev/null
for c in 1 3 4; do python tools/benchsum.py < gpurun_out/${TAG}_bench_C${c}_n1.json; done
python -c "
import json
for l in open('gpurun_out/${TAG}_bench_C5_n1.json'):
    if l.startswith('{'):
        d=json.loads(l); print('C5', round(d['value']), d['ms_per_step'], d.get('mutations_per_s'), d.get('more_chains',{}).get('value'))
"
echo "== done after $(( $(date +%s) - T0 )) s"
