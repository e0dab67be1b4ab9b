#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest bf16/persistent"; timeout 1200 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider -k "persistent or bf16 or smoke" > gpurun_out/pytest2.log 2>&1; tail -30 gpurun_out/pytest2.log
echo "=== smoke"; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke2.log 2>&1; tail -5 gpurun_out/smoke2.log
for mode in persistent streamed; do
echo "=== bench bf16 cfg3 $mode"; timeout 900 python bench.py --workload cfg3 --precision bf16 --rnn-mode $mode --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg3_bf16_$mode.log 2>&1; tail -2 gpurun_out/bench_cfg3_bf16_$mode.log | cut -c1-1800
done
echo "=== bench bf16 cfg2 persistent"; timeout 600 python bench.py --workload cfg2 --precision bf16 --rnn-mode persistent --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg2_bf16_persistent.log 2>&1; tail -2 gpurun_out/bench_cfg2_bf16_persistent.log | cut -c1-1800
