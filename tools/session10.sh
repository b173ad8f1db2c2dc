#!/bin/bash
# Round 2, GPU session 10 (8 GPUs): the scaling run the driver does at round end, rehearsed — N = 8 and 4 on C2 (with the secondary
# configs: C4 tile-split), the reference arm under torchrun, the multi-device tests on real peers.
T0=$(date +%s)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8; nproc
timeout 600 python -m pytest tests/test_multi_cabi.py tests/test_gpu_multi.py -q -m gpu -x 2>&1 | tail -3
echo "== tests done after $(( $(date +%s) - T0 )) s"
for N in 8 4; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520 + N)) bench.py --gpus $N --steps 3 --warmup 3 \
    > gpurun_out/r02_bench_C2_n$N.json 2> gpurun_out/r02_bench_C2_n$N.err; echo "N=$N rc=$?"; tail -2 gpurun_out/r02_bench_C2_n$N.err
  python tools/benchsum.py < gpurun_out/r02_bench_C2_n$N.json
  echo "== N=$N done after $(( $(date +%s) - T0 )) s"
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29540 bench.py --gpus 8 --workload C4 --steps 1 --warmup 1 --no-configs \
  > gpurun_out/r02_bench_C4_n8.json 2> gpurun_out/r02_bench_C4_n8.err; echo "C4 N=8 rc=$?"; python tools/benchsum.py < gpurun_out/r02_bench_C4_n8.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 \
  > gpurun_out/r02_bench_C2_reference_n8.json 2>/dev/null; echo "ref rc=$?"; cut -c1-400 gpurun_out/r02_bench_C2_reference_n8.json
echo "== done after $(( $(date +%s) - T0 )) s"
