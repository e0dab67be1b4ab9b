#!/bin/bash
mkdir -p gpurun_out
for rot in 0 1; do for hs in 16 32; do
echo "=== bench bf16 cfg3 persistent HS=$hs ROT=$rot"; MVAE_REC_ROT=$rot MVAE_REC_HS=$hs timeout 900 python bench.py --workload cfg3 --precision bf16 --rnn-mode persistent --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/b6_${hs}_${rot}.log 2>&1; tail -3 gpurun_out/b6_${hs}_${rot}.log | cut -c1-400
done; done
