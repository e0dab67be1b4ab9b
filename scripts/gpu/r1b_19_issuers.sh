#!/bin/bash
mkdir -p gpurun_out
echo "=== parity"
MVAE_CL_STAGGER=2000 MVAE_CLB_STAGGER=4000 timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "persistent_rnn or bf16" -p no:cacheprovider 2>&1 | grep "assert\|Error\|passed\|failed\|FAILED\|timeout\|trap" | head -20
for cfg in "0 0" "2500 0" "0 5000" "2500 5000" "1200 2500"; do
set -- $cfg
echo "=== stagger fwd=$1 bwd=$2"
MVAE_CL_STAGGER=$1 MVAE_CLB_STAGGER=$2 timeout 600 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['class_ms'])"
done
