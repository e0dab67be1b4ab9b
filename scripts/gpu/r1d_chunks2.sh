#!/bin/bash
# time-chunked layer pipeline: parity of the new paths, A/B of the chunk counts at cfg3, then the banked deliverables of the default build
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
show='import sys,json; d=json.loads(sys.stdin.read()); print(round(d["value"]), "seq/s", round(d["ms_per_step"],3), "ms", d["roofline"]["class_ms"], "loss", d["config"].get("loss_last_step"))'
for cfg in "2 1 0" "4 1 0" "4 1 1" "8 1 0" "4 2 0" "4 4 0"; do
  set -- $cfg
  echo "=== cfg3 chunks fwd $1 bwd $2 branch_at $3"
  MVAE_CHUNKS=$1 MVAE_CHUNKS_BWD=$2 MVAE_BRANCH_AT=$3 timeout 300 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r1d_bench_$1_$2_$3.log 2> gpurun_out/r1d_bench_$1_$2_$3.err
  tail -1 gpurun_out/r1d_bench_$1_$2_$3.log | python -c "$show" || tail -3 gpurun_out/r1d_bench_$1_$2_$3.err
done
echo "=== timeline chunks 4 1 0"
MVAE_TIMELINE=1 MVAE_CHUNKS=4 timeout 300 python bench.py --workload cfg3 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r1d_timeline_4.log 2> gpurun_out/r1d_timeline_4.err
grep timeline gpurun_out/r1d_timeline_4.err | tail -70 > gpurun_out/r1d_timeline_4.txt; wc -l gpurun_out/r1d_timeline_4.txt
echo "=== cfg2 chunks 1 / 2"
for c in 1 2; do
  MVAE_CHUNKS=$c timeout 300 python bench.py --workload cfg2 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2> gpurun_out/r1d_cfg2_$c.err | tail -1 | tee gpurun_out/r1d_cfg2_$c.log | python -c "$show"
done
