#!/bin/bash
# GPU-side script: parity tests of the pipelined K1 kernel, then timing of the variants and one ncu capture.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mtm_tma.py -m gpu -x -q 2>&1 | tail -30 > gpurun_out/pytest_k1tma.log
{
  echo "== old kernel"; SPYB_MTM_NO_TMA=1 timeout 120 python tools/k1_time.py
  for nst in 8 10 12; do echo "== NST=$nst"; SPYB_MTM_NST=$nst timeout 120 python tools/k1_time.py; done
  echo "== dpss old"; SPYB_MTM_NO_TMA=1 timeout 120 python tools/k1_time.py --taper dpss --iters 5
  echo "== dpss NST=10"; timeout 120 python tools/k1_time.py --taper dpss --iters 5
} > gpurun_out/k1_time_r2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mtm_tma -s 3 -c 1 -o gpurun_out/prof_r2_k1tma -f \
    python tools/k1_time.py --iters 1 > gpurun_out/ncu_k1tma.log 2>&1
