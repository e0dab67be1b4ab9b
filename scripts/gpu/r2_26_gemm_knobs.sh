#!/bin/bash
# round 2, call 26: split-K granularity, L2 prefetch distance and side-stream grid of the weight-gradient GEMMs; clock64 trace of the reverse sweep
set -x
mkdir -p gpurun_out
line='import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print(sys.argv[1],round(d["ms_per_step"],3),round(d["value"]),d["roofline"]["class_ms"])'
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "$line" $tag; }
run base X=1
run mult4 MVAE_GEMM_SPLIT_MULT=4
run mult6 MVAE_GEMM_SPLIT_MULT=6
run pf8 MVAE_GEMM_PREFETCH=8
run pf16 MVAE_GEMM_PREFETCH=16
run mult4_pf8 MVAE_GEMM_SPLIT_MULT=4 MVAE_GEMM_PREFETCH=8
run sms100 MVAE_SIDE_SMS=100
run sms148 MVAE_SIDE_SMS=148
run sms64 MVAE_SIDE_SMS=64
run base X=1
MVAE_REC_TRACE=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2> gpurun_out/r2_26_rec_trace.txt
timeout 200 python -c "
from midi_vae_b200 import _lib
import sys
sys.exit(1 if _lib.load().mvae_selftest_gemm(0, 0) else 0)" 2>&1 | tail -3
