#!/bin/bash
# first contact of the cluster/DSMEM forward kernel: parity vs the streamed form, then timing at cfg3
mkdir -p gpurun_out
export MVAE_CL_VERBOSE=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "=== parity (nswap=0)"
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "persistent_rnn and (shape1 or shape5 or shape6)" --maxfail=3 -p no:cacheprovider 2>&1 | tail -15
rc=${PIPESTATUS[0]}
echo "=== parity (nswap=1)"
MVAE_CL_NSWAP=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "persistent_rnn and (shape1 or shape5 or shape6)" --maxfail=3 -p no:cacheprovider 2>&1 | tail -15
echo "=== bench cfg3 cluster fwd"
MVAE_REC_TRACE=1 timeout 600 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r1b_01_bench.log 2> gpurun_out/r1b_01_bench.err
tail -1 gpurun_out/r1b_01_bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['class_ms'])"
grep -A9 "rec trace fwd" gpurun_out/r1b_01_bench.err | head -40
grep "rec_cluster" gpurun_out/r1b_01_bench.err | head
echo "=== bench cfg3 old persistent fwd (MVAE_REC_CLUSTER=0)"
MVAE_REC_CLUSTER=0 timeout 600 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r1b_01_bench_old.log 2>&1
tail -1 gpurun_out/r1b_01_bench_old.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['class_ms'])"
