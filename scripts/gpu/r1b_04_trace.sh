#!/bin/bash
mkdir -p gpurun_out
for cfg in "1 1" "1 0" "0 1"; do
set -- $cfg
echo "=== push_mode=$1 spin=$2 bench cfg3"
MVAE_CL_PUSH=$1 MVAE_CL_SPIN=$2 MVAE_REC_TRACE=1 timeout 600 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r1b_04_bench_$1$2.log 2> gpurun_out/r1b_04_bench_$1$2.err
tail -1 gpurun_out/r1b_04_bench_$1$2.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['class_ms'])"
grep -A8 "rec trace fwd" gpurun_out/r1b_04_bench_$1$2.err | head -9
done
