#!/bin/bash
# first GPU contact: GEMM selftest (verbose), parity tests, smoke, short benches
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
echo "=== selftest" ; timeout 300 python -c "
import sys; sys.path.insert(0,'.')
from midi_vae_b200 import _lib
print('rc', _lib.load().mvae_selftest_gemm(0, 1))
" > gpurun_out/selftest.log 2>&1; tail -5 gpurun_out/selftest.log
echo "=== pytest"; timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/pytest.log 2>&1; tail -40 gpurun_out/pytest.log
echo "=== smoke"; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -5 gpurun_out/smoke.log
echo "=== bench fp32 cfg2"; timeout 600 python bench.py --workload cfg2 --precision fp32 --steps 5 --warmup 3 > gpurun_out/bench_cfg2_fp32.log 2>&1; tail -3 gpurun_out/bench_cfg2_fp32.log
echo "=== bench bf16 cfg2"; timeout 600 python bench.py --workload cfg2 --precision bf16 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg2_bf16.log 2>&1; tail -3 gpurun_out/bench_cfg2_bf16.log
echo "=== bench bf16 cfg3"; timeout 900 python bench.py --workload cfg3 --precision bf16 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg3_bf16.log 2>&1; tail -3 gpurun_out/bench_cfg3_bf16.log
