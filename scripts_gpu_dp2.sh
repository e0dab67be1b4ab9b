#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_dp2.log 2>&1; tail -3 gpurun_out/bench_dp2.log | cut -c1-1500
timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_dp1.log 2>&1; tail -1 gpurun_out/bench_dp1.log | cut -c1-300
