#!/bin/bash
# round 2, call 34: GRU cluster kernels: exchanges by DSMEM stores + cluster barriers instead of bulk copies + mbarriers
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gru.py tests/test_gpu_classifier.py -m gpu -q -x 2>&1 | tail -4
MVAE_GRU_ROWS=32 timeout 600 python -m pytest tests/test_gpu_gru.py -m gpu -q -x -k "cluster or default_shape" 2>&1 | tail -3
line='import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print(sys.argv[1],round(d["ms_per_step"],3),round(d["value"]),d["roofline"]["class_ms"])'
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --workload refdefault --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>>gpurun_out/r2_34_bench.err | python -c "$line" $tag; }
run xs1 MVAE_GRU_XCHG=1
run xs0 MVAE_GRU_XCHG=0
run xs1 MVAE_GRU_XCHG=1
run xs0 MVAE_GRU_XCHG=0
run xs1_rows32 MVAE_GRU_XCHG=1 MVAE_GRU_ROWS=32
tail -3 gpurun_out/r2_34_bench.err
