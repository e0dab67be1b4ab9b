#!/bin/bash
mkdir -p gpurun_out
echo "=== full gpu tests"
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -8
echo "=== default bench"
timeout 900 python bench.py > gpurun_out/r1b_13_bench_default.log 2> gpurun_out/r1b_13_bench_default.err
tail -1 gpurun_out/r1b_13_bench_default.log
echo "=== smoke"
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -4
