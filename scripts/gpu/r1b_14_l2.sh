#!/bin/bash
mkdir -p gpurun_out
for l2 in 1 0; do
echo "=== bwd via_l2=$l2 parity"
MVAE_CLB_L2=$l2 timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "persistent_rnn and (shape1 or shape5 or shape6)" -p no:cacheprovider 2>&1 | grep "assert\|Error\|passed\|failed\|FAILED\|timeout\|trap" | head -20
echo "=== bwd via_l2=$l2 bench cfg3"
MVAE_CLB_L2=$l2 MVAE_REC_TRACE=1 timeout 600 python bench.py --workload cfg3 --steps 8 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r1b_14_bench_$l2.log 2> gpurun_out/r1b_14_bench_$l2.err
tail -1 gpurun_out/r1b_14_bench_$l2.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['class_ms'])"
grep -A8 "rec trace bwd" gpurun_out/r1b_14_bench_$l2.err | head -9 | grep "cta . step 1[78]"
done
