#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider 2>&1 | tail -4
for w in 1 0; do
echo "=== bench bf16 cfg3 BN256=$w"; MVAE_GEMM_BN256=$w timeout 900 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/b11_${w}.log 2>&1; tail -1 gpurun_out/b11_${w}.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['class_ms'])"
done
