#!/bin/bash
# round 2, call 24: column sums with 8 loads in flight per thread, on their own stream next to the weight-gradient GEMMs; velocity-head bias sum fused
set -x
mkdir -p gpurun_out
line='import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print(sys.argv[1],round(d["ms_per_step"],3),round(d["value"]),d["roofline"]["class_ms"])'
for rep in 1 2; do
  for v in 0 1; do
    MVAE_COLSUM_STREAM=$v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "$line" colsum_stream$v
  done
done
MVAE_GEMM_PAIR=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "$line" pair0
MVAE_TIMELINE=2 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2> gpurun_out/r2_24_timeline_cs1.txt
MVAE_COLSUM_STREAM=0 MVAE_TIMELINE=2 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2> gpurun_out/r2_24_timeline_cs0.txt
timeout 300 python bench.py --workload cfg2 --steps 50 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "$line" cfg2
timeout 600 python -m pytest tests -m gpu -q -x --durations=3 2>&1 | tail -8 > gpurun_out/r2_24_pytest.log; tail -4 gpurun_out/r2_24_pytest.log
