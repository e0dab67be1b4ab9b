#!/bin/bash
# round 2, call 28: cluster-resident GRU kernels (gru_cluster.cu): GRU GPU tests, reference-default bench line, streamed A/B
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gru.py tests/test_gpu_classifier.py -m gpu -q -x --durations=5 2>&1 | tail -30 > gpurun_out/r2_28_pytest_gru.log; tail -30 gpurun_out/r2_28_pytest_gru.log
line='import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print(sys.argv[1],round(d["ms_per_step"],3),round(d["value"]),d["roofline"]["class_ms"])'
timeout 300 python bench.py --workload refdefault --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>gpurun_out/r2_28_bench.err | python -c "$line" gru_cluster
MVAE_GRU_CLUSTER=0 timeout 300 python bench.py --workload refdefault --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>>gpurun_out/r2_28_bench.err | python -c "$line" gru_streamed
tail -5 gpurun_out/r2_28_bench.err
