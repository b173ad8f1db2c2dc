#!/bin/bash
# Round 2, GPU session 25 (2 GPUs): the ordered wavefront behind torchrun + NCCL and behind bn_render_multi, on real peers
mkdir -p gpurun_out
T0=$(date +%s)
timeout 600 python -m pytest tests/test_multi_cabi.py tests/test_gpu_multi.py -q -m gpu -x 2>&1 | tail -3
echo "== tests done after $(( $(date +%s) - T0 )) s"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 3 --warmup 3 \
  > gpurun_out/r02f_bench_C2_n2.json 2> gpurun_out/r02f_bench_C2_n2.err; echo "N=2 rc=$?"; tail -2 gpurun_out/r02f_bench_C2_n2.err
python tools/benchsum.py < gpurun_out/r02f_bench_C2_n2.json
echo "== done after $(( $(date +%s) - T0 )) s"
