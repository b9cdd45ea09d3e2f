#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mtm_tma.py -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_k1tma.log
{
  for dbg in 0 1 5; do echo "== DBG=$dbg"; SPYB_MTM_DBG=$dbg timeout 120 python tools/k1_time.py; done
  echo "== old"; SPYB_MTM_NO_TMA=1 timeout 120 python tools/k1_time.py
} > gpurun_out/k1_time_r2b.log 2>&1
