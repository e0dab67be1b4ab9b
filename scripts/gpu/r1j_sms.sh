#!/bin/bash
mkdir -p gpurun_out
show='import sys,json; d=json.loads(sys.stdin.read()); print(round(d["value"]), "seq/s", round(d["ms_per_step"],3), "ms", d["roofline"]["class_ms"])'
for sms in 84 100 116 132 148 68; do
  echo "=== cfg3 side_sms $sms"
  MVAE_SIDE_SMS=$sms timeout 300 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r1j_bench_sms$sms.log 2> gpurun_out/r1j_bench_sms$sms.err
  tail -1 gpurun_out/r1j_bench_sms$sms.log | python -c "$show" || tail -3 gpurun_out/r1j_bench_sms$sms.err
done
echo "=== cfg3 as_wired (the reference's own decoder wiring)"
timeout 300 python bench.py --workload cfg3 --feedback as_wired --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r1j_bench_as_wired.log 2> gpurun_out/r1j_bench_as_wired.err
tail -1 gpurun_out/r1j_bench_as_wired.log | python -c "$show" || tail -3 gpurun_out/r1j_bench_as_wired.err
