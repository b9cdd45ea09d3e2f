#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_v4_hann.json 2> gpurun_out/bench_r2_v4_hann.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2_v4_reference.json 2> gpurun_out/bench_r2_v4_reference.err
timeout 600 python bench.py --steps 10 --warmup 3 --taper dpss --no-configs > gpurun_out/bench_r2_v4_dpss.json 2> gpurun_out/bench_r2_v4_dpss.err
