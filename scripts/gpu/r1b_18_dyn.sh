#!/bin/bash
mkdir -p gpurun_out
echo "=== gemm selftest + parity"
timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "gemm or persistent_rnn or bf16" -p no:cacheprovider 2>&1 | grep "assert\|Error\|passed\|failed\|FAILED\|timeout\|trap" | head -20
echo "=== bench cfg3"
MVAE_TIMELINE=1 timeout 600 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r1b_18_bench.log 2> gpurun_out/r1b_18_bench.err
tail -1 gpurun_out/r1b_18_bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['class_ms'])"
grep timeline gpurun_out/r1b_18_bench.err | grep "gemm" | awk '$6>0.2'
