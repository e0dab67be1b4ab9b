#!/bin/bash
mkdir -p gpurun_out
echo "=== N=1 (bench.py after the roofline.traffic change)"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r1i_bench_n1.log 2> gpurun_out/r1i_bench_n1.err; tail -1 gpurun_out/r1i_bench_n1.log | cut -c1-160
python -c "import json; d=json.loads(open('gpurun_out/r1i_bench_n1.log').read().strip().splitlines()[-1]); print(d['roofline']['traffic'], d['roofline']['frac'], d['roofline']['step']['frac'])"
echo "=== N=2 data parallel"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r1i_bench_n2.log 2> gpurun_out/r1i_bench_n2.err
tail -1 gpurun_out/r1i_bench_n2.log | cut -c1-400; tail -3 gpurun_out/r1i_bench_n2.err | cut -c1-200
