#!/bin/bash
# Round 2, GPU session 34: what the driver runs at round end, on the final tree: pytest -m gpu, smoke(), the default bench line, the reference arm
mkdir -p gpurun_out
T0=$(date +%s)
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3 | tee gpurun_out/r02h_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== tests + smoke done after $(( $(date +%s) - T0 )) s"
timeout 600 python bench.py > gpurun_out/r02h_bench_default.json 2> gpurun_out/r02h_bench_default.err; echo "bench rc=$?"; tail -2 gpurun_out/r02h_bench_default.err
python tools/benchsum.py < gpurun_out/r02h_bench_default.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02h_bench_reference.json 2>/dev/null; echo "ref rc=$?"; cut -c1-300 gpurun_out/r02h_bench_reference.json
echo "== done after $(( $(date +%s) - T0 )) s"
