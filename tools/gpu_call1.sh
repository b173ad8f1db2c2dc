#!/bin/bash
# scratch GPU call: new parity tests, PSSMLT fast-trace A/B, phase-T repeat A/B, traversal phase stats
mkdir -p gpurun_out
{
echo "== pytest (published-image pin + PSSMLT parity with trace_lane)"
timeout 400 python -m pytest tests/test_ref_sample_image.py tests/test_pssmlt.py -q -m gpu 2>&1 | tail -15
echo "== MLT A/B"
for v in libbarnacle_b200.so lib_mltexact.so; do echo "-- $v"; BN_LIB=$PWD/barnacle_b200/lib/$v timeout 200 python tools/mlt_bench.py 2>&1 | tail -3; done
echo "== traversal A/B"
AB_WORKLOADS="C2:32 C4:4 C3:16" tools/ab.sh libbarnacle_b200.so lib_stayT12.so lib_stayT6.so
echo "== stats"
for s in "cbox_bunny 1024 1024 4" "bunny_instanced 3840 2160 1" "material_sweep 1920 1080 4"; do echo "-- $s"; BN_LIB=$PWD/barnacle_b200/lib/lib_stats.so timeout 120 python tools/trav_stats.py $s 2>&1 | tail -14; done
} > gpurun_out/call1.log 2>&1
tail -80 gpurun_out/call1.log
