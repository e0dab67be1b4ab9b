#!/bin/bash
# round 2, call 32 (2 GPUs): on-hardware data-parallel equivalence test and N = 2 bench lines of the final build (cfg3 LSTM, reference-default GRU)
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_dp.py -m gpu -q -s 2>&1 | tail -15 > gpurun_out/r2_32_pytest_dp.log; tail -6 gpurun_out/r2_32_pytest_dp.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_32_bench_n2.json 2> gpurun_out/r2_32_bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload refdefault --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_32_bench_n2_gru.json 2> gpurun_out/r2_32_bench_n2_gru.err
for f in n2 n2_gru; do python -c "
import json
d=json.loads(open('gpurun_out/r2_32_bench_$f.json').read().strip().splitlines()[-1])
print('$f', d['n_gpus'], d['ms_per_step'], d['value'], (d.get('e2e') or {}).get('value'), d['roofline'].get('class_ms', {}).get('allreduce'))"; done
tail -3 gpurun_out/r2_32_bench_n2.err gpurun_out/r2_32_bench_n2_gru.err
