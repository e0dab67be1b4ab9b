#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest persistent"; timeout 1200 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider -k "persistent or bf16" > gpurun_out/pytest3.log 2>&1; tail -5 gpurun_out/pytest3.log
for hs in 16 32; do
echo "=== bench bf16 cfg3 persistent HS=$hs"; MVAE_REC_HS=$hs timeout 900 python bench.py --workload cfg3 --precision bf16 --rnn-mode persistent --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_cfg3_hs$hs.log 2>&1; tail -1 gpurun_out/bench_cfg3_hs$hs.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['class_ms'])"
done
