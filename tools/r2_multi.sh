#!/bin/bash
# GPU-side (2+ GPUs): multi-GPU parity script + bench under torchrun
cd "$(dirname "$0")/.."
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tests/multi_gpu_check.py > gpurun_out/multi_gpu_check_n$N.log 2>&1
echo "multi_gpu_check rc=$?" >> gpurun_out/multi_gpu_check_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_r2_v8_n$N.json 2> gpurun_out/bench_r2_v8_n$N.err
echo "bench rc=$?" >> gpurun_out/multi_gpu_check_n$N.log
tail -3 gpurun_out/bench_r2_v8_n$N.err
