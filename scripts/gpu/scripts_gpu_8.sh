#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider -k "persistent or bf16" 2>&1 | tail -3
for hs in 16 32; do
echo "=== bench bf16 cfg3 persistent HS=$hs"; MVAE_REC_HS=$hs timeout 900 python bench.py --workload cfg3 --precision bf16 --rnn-mode persistent --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/b8_${hs}.log 2>&1; tail -1 gpurun_out/b8_${hs}.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['class_ms'])"
done
MVAE_REC_HS=16 MVAE_REC_TRACE=1 timeout 300 python scripts_one_step.py persistent 1 > gpurun_out/trace8.log 2>&1
grep -A2 "rec trace" gpurun_out/trace8.log | grep -A2 -E "rec trace (fwd lstm_1|bwd lstm_1|bwd notes/cell_2)" | cut -c1-330
