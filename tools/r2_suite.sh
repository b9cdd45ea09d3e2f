#!/bin/bash
# GPU-side: full parity suite + a short bench run (N = 1)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -40 > gpurun_out/pytest_gpu_r2_v1.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2_v1.json 2> gpurun_out/bench_r2_v1.err
tail -5 gpurun_out/bench_r2_v1.err
