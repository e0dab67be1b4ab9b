#!/bin/bash
mkdir -p gpurun_out
for pm in 2 0 1; do
echo "=== push_mode=$pm parity"
MVAE_CL_PUSH=$pm timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "persistent_rnn and (shape1 or shape5 or shape6)" -p no:cacheprovider 2>&1 | grep "assert\|Error\|passed\|failed\|FAILED\|timeout\|trap" | head -10
echo "=== push_mode=$pm bench cfg3"
MVAE_CL_PUSH=$pm MVAE_REC_TRACE=1 timeout 600 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r1b_05_bench_$pm.log 2> gpurun_out/r1b_05_bench_$pm.err
tail -1 gpurun_out/r1b_05_bench_$pm.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['class_ms'])"
grep -A8 "rec trace fwd" gpurun_out/r1b_05_bench_$pm.err | head -9
done
