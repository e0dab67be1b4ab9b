#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider -k "persistent" 2>&1 | tail -3
for hs in 16 32; do
MVAE_REC_HS=$hs MVAE_REC_TRACE=1 timeout 300 python scripts_one_step.py persistent 1 > gpurun_out/trace7_$hs.log 2>&1
grep -A2 "rec trace" gpurun_out/trace7_$hs.log | grep -E "rec trace (fwd lstm_1|bwd lstm_1|bwd notes/cell_2)|step" | cut -c1-330 | head -12
done
