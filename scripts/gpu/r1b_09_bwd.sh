#!/bin/bash
mkdir -p gpurun_out
export MVAE_CL_VERBOSE=1
echo "=== cluster bwd parity"
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "persistent_rnn and (shape1 or shape5 or shape6)" -p no:cacheprovider 2>&1 | grep "assert\|Error\|passed\|failed\|FAILED\|timeout\|trap\|rec_cluster" | head -20
for ng in 2 1; do
echo "=== bwd ng=$ng bench cfg3"
MVAE_CLB_NG=$ng MVAE_REC_TRACE=1 timeout 600 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r1b_09_bench_$ng.log 2> gpurun_out/r1b_09_bench_$ng.err
tail -1 gpurun_out/r1b_09_bench_$ng.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['class_ms'])"
grep "rec_cluster_bwd\|timeout\|error" gpurun_out/r1b_09_bench_$ng.err | head -4
grep -A8 "rec trace bwd" gpurun_out/r1b_09_bench_$ng.err | head -9 | grep -v "step 19\|step 20"
done
