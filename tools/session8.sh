#!/bin/bash
# Round 2, GPU session 8: PSSMLT as a wavefront over chains — parity (bootstrap weights, B, per-chain accepted counts, rays) and C5.
T0=$(date +%s)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pssmlt.py tests/test_ref_sample_image.py -q -m gpu -x 2>&1 | tail -15
echo "== tests done after $(( $(date +%s) - T0 )) s"
echo "== wavefront"; timeout 300 python bench.py --workload C5 --steps 3 --warmup 2 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('value', round(d['value']), 'Mrays/s  ms/step', round(d['ms_per_step'], 2), 'mut/s', round(d['mutations_per_s'] / 1e6, 1), 'M  acc', round(d['acceptance_rate'], 4), 'e2e', round(d['e2e']['value']))"
echo "== megakernel"; BN_MLT_MEGAKERNEL=1 timeout 300 python bench.py --workload C5 --steps 3 --warmup 2 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('value', round(d['value']), 'Mrays/s  ms/step', round(d['ms_per_step'], 2), 'mut/s', round(d['mutations_per_s'] / 1e6, 1), 'M  acc', round(d['acceptance_rate'], 4), 'e2e', round(d['e2e']['value']))"
timeout 120 python tools/mlt_bench.py 2>&1 | tail -12
echo "== done after $(( $(date +%s) - T0 )) s"
