#!/bin/bash
# round 2, call 2: whole GPU suite without -x (every failure listed), then compute-sanitizer on small cluster-kernel cases
set -x
mkdir -p gpurun_out
rm -f gpurun_out/parity_bench_shapes.jsonl
python -m pytest tests -m gpu -q -s --durations=15 2>&1 | tail -150 > gpurun_out/r2_02_pytest.log
tail -8 gpurun_out/r2_02_pytest.log
python scripts/sanitize_case.py 512 8 72 > gpurun_out/r2_02_plain_512.log 2>&1; tail -1 gpurun_out/r2_02_plain_512.log
for tool in memcheck synccheck racecheck; do
  for H in 512 256; do
    timeout 600 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_case.py $H 8 72 > gpurun_out/r2_02_san_${tool}_${H}.log 2>&1
    echo "$tool $H rc=$?"; tail -4 gpurun_out/r2_02_san_${tool}_${H}.log
  done
done
