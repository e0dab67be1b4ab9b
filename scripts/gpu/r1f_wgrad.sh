#!/bin/bash
# chunked reverse sweeps with per-chunk weight-gradient GEMMs: parity, then A/B of chunk count x side-stream grid at cfg3
mkdir -p gpurun_out
echo "=== parity (chunked pipeline)"
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "time_chunked" -p no:cacheprovider 2>&1 | tail -6
show='import sys,json; d=json.loads(sys.stdin.read()); print(round(d["value"]), "seq/s", round(d["ms_per_step"],3), "ms", d["roofline"]["class_ms"])'
for cfg in "1 0" "1 36" "2 0" "4 0" "4 36" "4 52" "8 0" "8 36"; do
  set -- $cfg
  echo "=== cfg3 chunks bwd $1 side_sms $2"
  MVAE_CHUNKS_BWD=$1 MVAE_SIDE_SMS=$2 timeout 300 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r1f_bench_$1_$2.log 2> gpurun_out/r1f_bench_$1_$2.err
  tail -1 gpurun_out/r1f_bench_$1_$2.log | python -c "$show" || tail -3 gpurun_out/r1f_bench_$1_$2.err
done
echo "=== timeline bwd 4"
MVAE_TIMELINE=1 MVAE_CHUNKS_BWD=4 timeout 300 python bench.py --workload cfg3 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r1f_timeline.log 2> gpurun_out/r1f_timeline.err
grep timeline gpurun_out/r1f_timeline.err | tail -110 > gpurun_out/r1f_timeline.txt; wc -l gpurun_out/r1f_timeline.txt
