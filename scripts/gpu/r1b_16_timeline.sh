#!/bin/bash
mkdir -p gpurun_out
MVAE_TIMELINE=1 timeout 600 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r1b_16_bench.log 2> gpurun_out/r1b_16_bench.err
tail -1 gpurun_out/r1b_16_bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['class_ms'])"
grep timeline gpurun_out/r1b_16_bench.err
