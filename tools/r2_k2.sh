#!/bin/bash
cd "$(dirname "$0")/.."
{
echo "== default"; timeout 100 python tools/k2f_profile.py; timeout 100 python tools/k2_accuracy.py
echo "== REWRITE_HI=0"; SPYB_TC_REWRITE_HI=0 timeout 100 python tools/k2f_profile.py; SPYB_TC_REWRITE_HI=0 timeout 100 python tools/k2_accuracy.py
echo "== CHAIN 128"; SPYB_TC_CHAIN_ROWS=128 timeout 100 python tools/k2f_profile.py; SPYB_TC_CHAIN_ROWS=128 timeout 100 python tools/k2_accuracy.py
echo "== CHAIN 32"; SPYB_TC_CHAIN_ROWS=32 timeout 100 python tools/k2f_profile.py
echo "== rows 1400"; ROWS=1400 timeout 100 python tools/k2f_profile.py
} > gpurun_out/k2_variants.log 2>&1
