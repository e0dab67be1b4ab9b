#!/bin/bash
# round 2, call 10 (8 GPUs): weak scaling at N = 8 with the gradient all-reduce in one piece vs two buckets, N = 1 on the same box beside it
set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r2_10_bench_n1.json 2> gpurun_out/r2_10_bench_n1.err
for bk in 2 1 2 1; do
  MVAE_AR_BUCKETS=$bk python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r2_10_bench_n8_buckets${bk}.json 2> gpurun_out/r2_10_bench_n8_buckets${bk}.err
  python -c "
import json
d=json.loads(open('gpurun_out/r2_10_bench_n8_buckets${bk}.json').read().strip().splitlines()[-1])
print('n8 buckets', $bk, d['ms_per_step'], d['value'], d['roofline'].get('class_ms', {}).get('allreduce'), d['clocks'])"
done
python -c "
import json
d=json.loads(open('gpurun_out/r2_10_bench_n1.json').read().strip().splitlines()[-1])
print('n1', d['ms_per_step'], d['value'], d['clocks'])"
