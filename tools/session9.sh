#!/bin/bash
# Round 2, GPU session 9: cooperative candidate pass (two lanes per new ray) vs the previous commit; PSSMLT forms.
T0=$(date +%s)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
echo "== tests done after $(( $(date +%s) - T0 )) s"
AB_WORKLOADS="C1:64 C2:32 C4:4 C3:16" tools/ab.sh lib_prev.so libbarnacle_b200.so lib_prev.so libbarnacle_b200.so
echo "== A/B done after $(( $(date +%s) - T0 )) s"
timeout 300 python bench.py --workload C5 --steps 3 --warmup 2 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('C5 value', round(d['value']), 'Mrays/s  ms/step', round(d['ms_per_step'], 2), 'mut/s', round(d['mutations_per_s'] / 1e6, 1), 'M e2e', round(d['e2e']['value'])); print('more_chains', d['more_chains'])"
echo "== done after $(( $(date +%s) - T0 )) s"
