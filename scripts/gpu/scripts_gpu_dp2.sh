#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
NCCL_DEBUG=WARN TORCH_NCCL_HEARTBEAT_TIMEOUT_SEC=120 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_dp2.log 2>&1; tail -1 gpurun_out/bench_dp2.log | cut -c1-2500
timeout 200 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_dp1.log 2>&1; tail -1 gpurun_out/bench_dp1.log | cut -c1-300
