#!/bin/bash
# multi-GPU evidence on one 8-GPU box: NCCL parity test, C2 strong scaling at N = 8 (and C4 tile/sample split)
N=${1:-8}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_multi.py -q -m gpu 2>&1 | tail -2 | tee gpurun_out/r01b_pytest_multi.txt
run() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" 2>gpurun_out/scale.err | grep '^{' ; }
run --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r01b_bench_C2_n$N.json
run --workload C4 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r01b_bench_C4_n$N.json
for c in 2 4; do python tools/benchsum.py < gpurun_out/r01b_bench_C${c}_n$N.json; done
tail -3 gpurun_out/scale.err
