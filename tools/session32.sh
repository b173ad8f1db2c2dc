#!/bin/bash
# Round 2, GPU session 32 (8 GPUs): the round's final kernels behind torchrun + NCCL and behind ONE bn_render_multi call, C2 at full size
T0=$(date +%s)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8; nproc
timeout 600 python -m pytest tests/test_multi_cabi.py tests/test_gpu_multi.py -q -m gpu -x 2>&1 | tail -3
echo "== tests done after $(( $(date +%s) - T0 )) s"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29528 bench.py --gpus 8 --steps 3 --warmup 3 --no-configs \
  > gpurun_out/r02g_bench_C2_n8.json 2> gpurun_out/r02g_bench_C2_n8.err; echo "N=8 rc=$?"; tail -2 gpurun_out/r02g_bench_C2_n8.err
python tools/benchsum.py < gpurun_out/r02g_bench_C2_n8.json
echo "== done after $(( $(date +%s) - T0 )) s"
