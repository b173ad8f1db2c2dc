#!/bin/bash
# Round 2, GPU session 5 (2 GPUs): the multi-GPU paths — torchrun bench (NCCL reduce; e2e through bn_render_multi from rank 0),
# the reference arm under torchrun, the 2-device tests.
T0=$(date +%s)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader; nproc
timeout 600 python -m pytest tests/test_multi_cabi.py tests/test_gpu_multi.py -q -m gpu -x 2>&1 | tail -3
echo "== tests done after $(( $(date +%s) - T0 )) s"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 \
  > gpurun_out/r02d_bench_C2_n2.json 2> gpurun_out/r02d_bench_C2_n2.err; echo "rc=$?"; tail -3 gpurun_out/r02d_bench_C2_n2.err
python tools/benchsum.py < gpurun_out/r02d_bench_C2_n2.json
echo "== bench N=2 done after $(( $(date +%s) - T0 )) s"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload C4 --spp 16 --steps 2 --warmup 1 --no-configs \
  > gpurun_out/r02d_bench_C4_n2.json 2> gpurun_out/r02d_bench_C4_n2.err; echo "rc=$?"; tail -3 gpurun_out/r02d_bench_C4_n2.err
python tools/benchsum.py < gpurun_out/r02d_bench_C4_n2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 \
  > gpurun_out/r02d_bench_C2_reference_n2.json 2>/dev/null; echo "rc=$?"; cat gpurun_out/r02d_bench_C2_reference_n2.json | cut -c1-600
echo "== done after $(( $(date +%s) - T0 )) s"
