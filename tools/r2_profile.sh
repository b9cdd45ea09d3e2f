#!/bin/bash
# GPU-side: full parity suite, smoke, bench (N = 1), ncu launch list + full captures of the two kernels of a step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu_r2_v2.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_r2.log 2>&1
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_v2_hann.json 2> gpurun_out/bench_r2_v2_hann.err
timeout 600 python bench.py --steps 10 --warmup 3 --taper dpss --no-configs > gpurun_out/bench_r2_v2_dpss.json 2> gpurun_out/bench_r2_v2_dpss.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2_v2.csv \
    python bench.py --steps 2 --warmup 1 --no-configs --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mtm_tma|csd_tc" -s 4 -c 2 -o gpurun_out/prof_r2_v2_hann -f \
    python tools/profile_step.py --iters 4 > gpurun_out/ncu_r2_v2.log 2>&1
