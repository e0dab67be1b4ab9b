#!/bin/bash
# round 2, call 37: GRU cluster path with the weight-gradient GEMMs on the side stream (as on the LSTM cluster path)
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gru.py tests/test_gpu_classifier.py -m gpu -q -x 2>&1 | tail -3
line='import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print(sys.argv[1],round(d["ms_per_step"],3),round(d["value"]),d["roofline"]["class_ms"])'
run() { tag=$1; shift; env "$@" timeout 200 python bench.py --workload refdefault --steps 30 --warmup 5 --no-cpu-baseline --no-e2e 2>>gpurun_out/r2_37_bench.err | python -c "$line" $tag; }
run side1 X=1
run side0 MVAE_SIDE_STREAM=0
run side1 X=1
run side0 MVAE_SIDE_STREAM=0
tail -n 3 gpurun_out/r2_37_bench.err
