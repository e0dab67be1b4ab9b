#!/bin/bash
mkdir -p gpurun_out
echo "=== parity (bwd4)"
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "persistent_rnn and (shape5 or shape7)" -p no:cacheprovider 2>&1 | grep "assert\|Error\|passed\|failed\|FAILED\|timeout\|trap" | head -12
for v in 4 2; do
echo "=== bwd form $v bench cfg3"
MVAE_CLB_V=$v MVAE_REC_TRACE=1 timeout 600 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r1b_25_bench_$v.log 2> gpurun_out/r1b_25_bench_$v.err
tail -1 gpurun_out/r1b_25_bench_$v.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['class_ms'])"
grep -A8 "rec trace bwd" gpurun_out/r1b_25_bench_$v.err | head -9 | grep "cta 0 step 1[78]"
tail -2 gpurun_out/r1b_25_bench_$v.err | cut -c1-300
done
