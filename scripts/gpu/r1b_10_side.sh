#!/bin/bash
mkdir -p gpurun_out
echo "=== parity"
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "persistent_rnn and (shape1 or shape5 or shape6) or bf16" -p no:cacheprovider 2>&1 | grep "assert\|Error\|passed\|failed\|FAILED\|timeout\|trap" | head -20
for cfg in "1 0" "1 100" "1 64" "0 0"; do
set -- $cfg
echo "=== side=$1 sms=$2 bench cfg3"
MVAE_SIDE_STREAM=$1 MVAE_SIDE_SMS=$2 timeout 600 python bench.py --workload cfg3 --steps 8 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r1b_10_bench_$1_$2.log 2> gpurun_out/r1b_10_bench_$1_$2.err
tail -1 gpurun_out/r1b_10_bench_$1_$2.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['class_ms'])"
done
