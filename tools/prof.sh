#!/bin/bash
# ncu --set full capture of the traversal kernels (extend + shadow of bounces 1-2) for one build.
# usage: tools/prof.sh <lib .so name under barnacle_b200/lib> <tag> [kernel regex] [workload] [spp]
LIBN=${1:-libbarnacle_b200.so}; TAG=${2:-x}; KRE=${3:-^k_traverse$}; WL=${4:-C2}; SPP=${5:-8}
mkdir -p gpurun_out
BN_LIB=$PWD/barnacle_b200/lib/$LIBN timeout ${PROF_TIMEOUT:-280} ncu --set full --clock-control none --import-source on -k regex:$KRE -s 2 -c 4 \
  -o gpurun_out/prof_$TAG -f python bench.py --workload $WL --spp $SPP --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/prof_$TAG.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/prof_$TAG.ncu-rep
