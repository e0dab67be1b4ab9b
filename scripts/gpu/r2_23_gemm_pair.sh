#!/bin/bash
# round 2, call 23: CTA-pair (cta_group::2, 256 x 256 tile) form of the tcgen05 GEMM: self test, then A/B on the cfg3 step
set -x
mkdir -p gpurun_out
timeout 200 python -c "
from midi_vae_b200 import _lib
import sys
sys.exit(1 if _lib.load().mvae_selftest_gemm(0, 0) else 0)" > gpurun_out/r2_23_selftest.log 2>&1
rc=$?; echo "selftest rc=$rc"; tail -12 gpurun_out/r2_23_selftest.log
if [ $rc -ne 0 ]; then exit 1; fi
line='import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print(sys.argv[1],round(d["ms_per_step"],3),round(d["value"]),d["roofline"]["class_ms"])'
for rep in 1 2; do
  for v in 0 1; do
    MVAE_GEMM_PAIR=$v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>gpurun_out/r2_23_bench_pair$v.err | python -c "$line" pair$v
  done
done
MVAE_GEMM_REDV4=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "$line" pair1_redv4_0
MVAE_GEMM_PAIR=0 MVAE_GEMM_REDV4=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "$line" pair0_redv4_0
MVAE_TIMELINE=2 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2> gpurun_out/r2_23_timeline_pair1.txt
MVAE_GEMM_PAIR=0 MVAE_TIMELINE=2 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2> gpurun_out/r2_23_timeline_pair0.txt
timeout 300 python bench.py --workload cfg2 --steps 50 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "$line" cfg2
timeout 300 python bench.py --workload cfg4 --steps 20 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "$line" cfg4
timeout 600 python -m pytest tests -m gpu -q -x --durations=3 2>&1 | tail -8 > gpurun_out/r2_23_pytest.log; tail -4 gpurun_out/r2_23_pytest.log
