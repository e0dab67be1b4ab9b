#!/bin/bash
# round 2, call 29: GRU cluster kernels with 16 rows per cluster and the independent recurrences on the branch stream
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gru.py tests/test_gpu_classifier.py -m gpu -q -x --durations=3 2>&1 | tail -12 > gpurun_out/r2_29_pytest_gru.log; tail -12 gpurun_out/r2_29_pytest_gru.log
line='import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print(sys.argv[1],round(d["ms_per_step"],3),round(d["value"]),d["roofline"]["class_ms"])'
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --workload refdefault --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>>gpurun_out/r2_29_bench.err | python -c "$line" $tag; }
run default X=1
run rows32 MVAE_GRU_ROWS=32
run nobranch MVAE_BRANCH=0
run rows32_nobranch MVAE_GRU_ROWS=32 MVAE_BRANCH=0
run nograph MVAE_STEP_GRAPH=0
run default X=1
MVAE_TIMELINE=1 timeout 300 python bench.py --workload refdefault --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2> gpurun_out/r2_29_timeline_gru.txt
tail -3 gpurun_out/r2_29_bench.err
