#!/bin/bash
# round 2, call 1: whole GPU suite (incl. the new bench-shape parity, GRU and facade tests), smoke, and the bench lines of every BASELINE config
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv
nproc
python -m pytest tests -m gpu -x -q -s --durations=15 2>&1 | tail -80 > gpurun_out/r2_01_pytest.log
tail -5 gpurun_out/r2_01_pytest.log
python __graft_entry__.py smoke > gpurun_out/r2_01_smoke.log 2>&1; tail -5 gpurun_out/r2_01_smoke.log
python bench.py > gpurun_out/r2_01_bench_cfg3.json 2> gpurun_out/r2_01_bench_cfg3.err; tail -c 600 gpurun_out/r2_01_bench_cfg3.json
python bench.py --workload cfg2 --steps 50 > gpurun_out/r2_01_bench_cfg2.json 2>> gpurun_out/r2_01_bench_cfg3.err
python bench.py --workload cfg4 --steps 20 > gpurun_out/r2_01_bench_cfg4.json 2>> gpurun_out/r2_01_bench_cfg3.err
python bench.py --workload cfg4 --cfg4-shape cfg2 --steps 20 > gpurun_out/r2_01_bench_cfg4_cfg2shape.json 2>> gpurun_out/r2_01_bench_cfg3.err
python bench.py --workload cfg5 > gpurun_out/r2_01_bench_cfg5.json 2>> gpurun_out/r2_01_bench_cfg3.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_01_bench_reference.json 2>> gpurun_out/r2_01_bench_cfg3.err
MVAE_TIMELINE=1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2> gpurun_out/r2_01_timeline.txt
tail -3 gpurun_out/r2_01_bench_cfg3.err
