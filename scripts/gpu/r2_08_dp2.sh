#!/bin/bash
# round 2, call 8 (2 GPUs): on-hardware data-parallel equivalence test, N = 1 / 2 bench with the gradient all-reduce in one piece vs two buckets
set -x
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests/test_gpu_dp.py -m gpu -q -s 2>&1 | tail -15 > gpurun_out/r2_08_pytest_dp.log; tail -6 gpurun_out/r2_08_pytest_dp.log
export MVAE_CLB_STM=1
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_08_bench_n1.json 2> gpurun_out/r2_08_bench_n1.err
for bk in 2 1; do
  MVAE_AR_BUCKETS=$bk python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_08_bench_n2_buckets${bk}.json 2> gpurun_out/r2_08_bench_n2_buckets${bk}.err
done
for f in n1 n2_buckets2 n2_buckets1; do python -c "
import json
d=json.loads(open('gpurun_out/r2_08_bench_$f.json').read().strip().splitlines()[-1])
print('$f', d['ms_per_step'], d['value'], d['roofline'].get('class_ms', {}).get('allreduce'))"; done
