#!/bin/bash
cd "$(dirname "$0")/.."
{
for p in 0 1 2 3; do echo "== L2PROMO=$p"; SPYB_MTM_L2PROMO=$p timeout 100 python tools/k1_time.py; done
echo "== STCS"; SPYB_MTM_DBG=16 timeout 100 python tools/k1_time.py
echo "== STCS promo0"; SPYB_MTM_DBG=16 SPYB_MTM_L2PROMO=0 timeout 100 python tools/k1_time.py
echo "== K2 chain128"; SPYB_TC_CHAIN_ROWS=128 timeout 100 python tools/k2f_profile.py; SPYB_TC_CHAIN_ROWS=128 timeout 100 python tools/k2_accuracy.py
echo "== K2 default"; timeout 100 python tools/k2f_profile.py
} > gpurun_out/k1_variants.log 2>&1
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:mtm_tma -s 3 -c 1 --csv --log-file gpurun_out/k1_promo0.csv env SPYB_MTM_L2PROMO=0 python tools/k1_time.py --iters 1 > /dev/null 2>&1
