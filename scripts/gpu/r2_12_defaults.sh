#!/bin/bash
# round 2, call 12: repeated A/B of the two schedule switches that looked positive (projection overlap, branch sweeps in waves of 3 clusters)
set -x
mkdir -p gpurun_out
show() { python -c "import json,sys;d=json.loads(open('$1').read().strip().splitlines()[-1]);print('$2',round(d['ms_per_step'],3),round(d['value']))"; }
for rep in 1 2 3; do
  for v in "MVAE_NONE=1" "MVAE_XW_OVERLAP=8" "MVAE_BRANCH_BWD_NCL=3" "MVAE_XW_OVERLAP=8 MVAE_BRANCH_BWD_NCL=3" "MVAE_XW_OVERLAP=16 MVAE_BRANCH_BWD_NCL=3"; do
    f=gpurun_out/r2_12_bench_$(echo $v | tr ' =' '__')_$rep.json
    env $v python bench.py --steps 15 --warmup 4 --no-cpu-baseline --no-e2e > $f 2>/dev/null; show $f "rep$rep $v"
  done
done
python -c "
from midi_vae_b200 import _lib
import sys
rc = _lib.load().mvae_selftest_gemm(0, 0); print('selftest rc', rc)" 2>&1 | tail -2
MVAE_XW_OVERLAP=8 MVAE_BRANCH_BWD_NCL=3 python -m pytest tests/test_gpu_parity_bench_shapes.py tests/test_gpu_parity.py -m gpu -q -x -k "cfg3 or cfg2 or persistent_rnn or bf16 or overlap" 2>&1 | tail -3
