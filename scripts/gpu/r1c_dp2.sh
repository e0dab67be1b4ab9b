#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r1c_bench_dp2.log 2> gpurun_out/r1c_bench_dp2.err
tail -1 gpurun_out/r1c_bench_dp2.log
tail -3 gpurun_out/r1c_bench_dp2.err
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r1c_bench_dp1.log 2>/dev/null
tail -1 gpurun_out/r1c_bench_dp1.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('dp1', d['value'], d['ms_per_step'])"
