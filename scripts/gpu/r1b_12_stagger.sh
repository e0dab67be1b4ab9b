#!/bin/bash
mkdir -p gpurun_out
echo "=== parity"
MVAE_CL_STAGGER=2500 MVAE_CLB_STAGGER=4000 timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "persistent_rnn and (shape1 or shape5 or shape6)" -p no:cacheprovider 2>&1 | grep "assert\|Error\|passed\|failed\|FAILED\|timeout\|trap" | head -20
for cfg in "0 0" "2500 4500" "1200 2500"; do
set -- $cfg
echo "=== stagger fwd=$1 bwd=$2 bench cfg3"
MVAE_CL_STAGGER=$1 MVAE_CLB_STAGGER=$2 MVAE_REC_TRACE=1 timeout 600 python bench.py --workload cfg3 --steps 8 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r1b_12_bench_$1.log 2> gpurun_out/r1b_12_bench_$1.err
tail -1 gpurun_out/r1b_12_bench_$1.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['class_ms'])"
grep -A8 "rec trace bwd" gpurun_out/r1b_12_bench_$1.err | head -9 | grep "cta 0 step 1[78]"
grep -A8 "rec trace fwd" gpurun_out/r1b_12_bench_$1.err | head -9 | grep "cta 0 step 1[78]"
done
