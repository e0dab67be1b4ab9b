#!/bin/bash
# round 2, call 36 (8 GPUs): weak scaling of the final build at N = 8 with the clock sampler started before the barrier
set -x
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_36_bench_n8.json 2> gpurun_out/r2_36_bench_n8.err
python -c "
import json
d=json.loads(open('gpurun_out/r2_36_bench_n8.json').read().strip().splitlines()[-1])
print('n8', d['n_gpus'], d['ms_per_step'], d['value'], (d.get('e2e') or {}).get('value'), d['roofline'].get('class_ms', {}).get('allreduce'), d['clocks'])"
tail -n 3 gpurun_out/r2_36_bench_n8.err
