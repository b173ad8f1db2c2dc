#!/bin/bash
# Round 2, GPU session 11: what ray ordering is worth to the traversal kernel (octant only / octant chunks as a producer could bin them / + Morton)
mkdir -p gpurun_out
for sc in cbox_bunny material_sweep; do echo "== $sc"; SORT_SPP=8 timeout 600 python tools/sort_experiment.py $sc; done
