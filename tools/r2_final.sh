#!/bin/bash
# GPU-side, end of round 2: full parity suite, smoke, default bench line, DPSS variant, reference arm (N = 1)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
V=${1:-v9}
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_gpu_r2_$V.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_r2_$V.log 2>&1
timeout 900 python bench.py > gpurun_out/bench_r2_${V}_hann.json 2> gpurun_out/bench_r2_${V}_hann.err
timeout 600 python bench.py --steps 10 --warmup 3 --taper dpss --no-configs > gpurun_out/bench_r2_${V}_dpss.json 2> gpurun_out/bench_r2_${V}_dpss.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2_${V}_reference.json 2> gpurun_out/bench_r2_${V}_reference.err
tail -3 gpurun_out/pytest_gpu_r2_$V.log; tail -2 gpurun_out/smoke_r2_$V.log; tail -2 gpurun_out/bench_r2_${V}_hann.err
