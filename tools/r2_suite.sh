#!/bin/bash
# GPU-side: full parity suite + smoke + a default bench run (N = 1)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
V=${1:-v7}
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_gpu_r2_$V.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_r2_$V.log 2>&1
timeout 900 python bench.py > gpurun_out/bench_r2_${V}_hann.json 2> gpurun_out/bench_r2_${V}_hann.err
tail -3 gpurun_out/pytest_gpu_r2_$V.log; tail -2 gpurun_out/smoke_r2_$V.log; tail -3 gpurun_out/bench_r2_${V}_hann.err
