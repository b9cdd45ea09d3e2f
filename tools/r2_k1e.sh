#!/bin/bash
# GPU-side: four-step K1 -- parity, timing against the 8-channel kernel, bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/k1_4s.log
: > $L
timeout 300 python -m pytest tests/test_gpu_mtm_4s.py -q -x 2>&1 | tail -15 >> $L
for M in 1 0; do
  echo "== SPYB_MTM_4S=$M" >> $L
  SPYB_MTM_4S=$M timeout 120 python tools/k1_time.py >> $L 2>&1
  SPYB_MTM_4S=$M timeout 120 python tools/k1_time.py --polyremoval -1 >> $L 2>&1
  SPYB_MTM_4S=$M SPYB_MTM_DBG=1 timeout 120 python tools/k1_time.py 2>&1 | sed 's/^/   stores off: /' >> $L
done
timeout 300 python -m pytest tests/test_gpu_csd_chains.py -q -x 2>&1 | tail -3 >> $L
cat $L
