#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider 2>&1 | tail -4
for pr in 1 0; do
echo "=== bench bf16 cfg3 PAIR=$pr"; MVAE_REC_PAIR=$pr timeout 900 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/b15_$pr.log 2>&1; tail -1 gpurun_out/b15_$pr.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['class_ms'])"
done
MVAE_REC_TRACE=1 timeout 300 python scripts_one_step.py persistent 1 > gpurun_out/trace15.log 2>&1
grep -A2 "rec trace" gpurun_out/trace15.log | grep -A2 -E "rec trace (bwd)" | cut -c1-330 | head -9
