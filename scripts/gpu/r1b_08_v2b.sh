#!/bin/bash
mkdir -p gpurun_out
echo "=== v2 parity (default ng)"
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "persistent_rnn and (shape1 or shape5 or shape6)" -p no:cacheprovider 2>&1 | grep "assert\|Error\|passed\|failed\|FAILED\|timeout\|trap" | head -10
for cfg in "2 0" "3 0" "3 1" "3 3" "1 0"; do
set -- $cfg
echo "=== v2 ng=$1 dbg=$2 bench cfg3"
MVAE_CL_NG=$1 MVAE_CL_DBG=$2 MVAE_REC_TRACE=1 timeout 600 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r1b_08_bench_$1$2.log 2> gpurun_out/r1b_08_bench_$1$2.err
tail -1 gpurun_out/r1b_08_bench_$1$2.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['class_ms'])"
grep -A8 "rec trace fwd" gpurun_out/r1b_08_bench_$1$2.err | head -9 | grep -v "step 19\|step 20"
done
