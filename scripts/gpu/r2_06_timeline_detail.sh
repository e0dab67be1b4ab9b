#!/bin/bash
# round 2, call 6: per-launch timeline of the weight-gradient stream; sanitizer probe variants (launch-attribute clusters, register-operand barriers)
set -x
mkdir -p gpurun_out
MVAE_CLB_STM=1 MVAE_TIMELINE=2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2> gpurun_out/r2_06_timeline_detail.txt
grep -c timeline gpurun_out/r2_06_timeline_detail.txt
for a in "barreg" "bulkx 8192 2 1" "bulkx 8192 8 1" "bulkx 8192 8 0" "bulkx 90000 16 5"; do
  for tool in memcheck synccheck; do
    timeout 120 compute-sanitizer --tool $tool --print-limit 3 scripts/sanitizer_probe $a > "gpurun_out/r2_06_probe_${tool}_${a// /_}.log" 2>&1
    echo "$tool $a rc=$?"; grep -E "probe|SUMMARY|Invalid|Barrier error|Race|not located|launch" "gpurun_out/r2_06_probe_${tool}_${a// /_}.log" | head -4
  done
done
