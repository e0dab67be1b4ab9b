#!/bin/bash
# projection overlap (stream-ordered waits on per-chunk counters of the forward cluster kernel): parity, then A/B at cfg3 / cfg2
mkdir -p gpurun_out
echo "=== parity (projection overlap)"
timeout 240 python -m pytest tests/test_gpu_parity.py -q -x -k "projection_overlap" -p no:cacheprovider 2>&1 | tail -8
show='import sys,json; d=json.loads(sys.stdin.read()); print(round(d["value"]), "seq/s", round(d["ms_per_step"],3), "ms", d["roofline"]["class_ms"])'
for nc in 0 2 4 8 16; do
  echo "=== cfg3 xw_overlap $nc"
  MVAE_XW_OVERLAP=$nc timeout 120 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r1k_bench_$nc.log 2> gpurun_out/r1k_bench_$nc.err
  tail -1 gpurun_out/r1k_bench_$nc.log | python -c "$show" || tail -3 gpurun_out/r1k_bench_$nc.err
done
echo "=== cfg2 xw_overlap 0 / 4"
for nc in 0 4; do MVAE_XW_OVERLAP=$nc timeout 120 python bench.py --workload cfg2 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2> gpurun_out/r1k_cfg2_$nc.err | tail -1 | tee gpurun_out/r1k_cfg2_$nc.log | python -c "$show"; done
echo "=== timeline overlap 4"
MVAE_TIMELINE=1 MVAE_XW_OVERLAP=4 timeout 120 python bench.py --workload cfg3 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r1k_timeline.log 2> gpurun_out/r1k_timeline.err
grep timeline gpurun_out/r1k_timeline.err | tail -61 > gpurun_out/r1k_timeline.txt; wc -l gpurun_out/r1k_timeline.txt
