#!/bin/bash
# round 2, call 25: weight-gradient launches deferred behind the next sweep's launch; branch-stream reverse sweeps as waves of 3 clusters
set -x
mkdir -p gpurun_out
line='import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print(sys.argv[1],round(d["ms_per_step"],3),round(d["value"]),d["roofline"]["class_ms"])'
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "$line" $tag; }
for rep in 1 2; do
  run d1w1 MVAE_WGRAD_DEFER=1 MVAE_BRANCH_WAVES=1
  run d0w0 MVAE_WGRAD_DEFER=0 MVAE_BRANCH_WAVES=0
  run d1w0 MVAE_WGRAD_DEFER=1 MVAE_BRANCH_WAVES=0
  run d0w1 MVAE_WGRAD_DEFER=0 MVAE_BRANCH_WAVES=1
done
MVAE_TIMELINE=2 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2> gpurun_out/r2_25_timeline_d1w1.txt
MVAE_WGRAD_DEFER=1 MVAE_BRANCH_WAVES=0 MVAE_TIMELINE=2 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2> gpurun_out/r2_25_timeline_d1w0.txt
MVAE_WGRAD_DEFER=0 MVAE_BRANCH_WAVES=1 MVAE_TIMELINE=2 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2> gpurun_out/r2_25_timeline_d0w1.txt
timeout 300 python bench.py --workload cfg2 --steps 50 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "$line" cfg2
timeout 600 python -m pytest tests -m gpu -q -x --durations=3 2>&1 | tail -8 > gpurun_out/r2_25_pytest.log; tail -4 gpurun_out/r2_25_pytest.log
