#!/bin/bash
# round 2, call 4: full GPU suite on the staged-epilogue GEMM + device post-processing, A/B of the staged epilogue at cfg3 / cfg2, sanitizer probes
set -x
mkdir -p gpurun_out
rm -f gpurun_out/parity_bench_shapes.jsonl
python -c "
from midi_vae_b200 import _lib
import sys
rc = _lib.load().mvae_selftest_gemm(0, 0); print('selftest rc', rc); sys.exit(rc)" > gpurun_out/r2_04_selftest.log 2>&1; tail -3 gpurun_out/r2_04_selftest.log
python -m pytest tests -m gpu -q -s --durations=8 2>&1 | tail -120 > gpurun_out/r2_04_pytest.log
tail -6 gpurun_out/r2_04_pytest.log
for so in 1 0; do
  MVAE_GEMM_STAGED_OUT=$so python bench.py --steps 15 --warmup 4 --no-cpu-baseline --no-e2e > gpurun_out/r2_04_bench_cfg3_staged${so}.json 2> gpurun_out/r2_04_bench_cfg3_staged${so}.err
  MVAE_GEMM_STAGED_OUT=$so python bench.py --workload cfg2 --steps 50 --no-cpu-baseline --no-e2e > gpurun_out/r2_04_bench_cfg2_staged${so}.json 2>> gpurun_out/r2_04_bench_cfg3_staged${so}.err
  MVAE_GEMM_STAGED_OUT=$so python bench.py --workload cfg4 --steps 20 --no-cpu-baseline --no-e2e > gpurun_out/r2_04_bench_cfg4_staged${so}.json 2>> gpurun_out/r2_04_bench_cfg3_staged${so}.err
  for w in cfg3 cfg2 cfg4; do python -c "import json;d=json.loads(open('gpurun_out/r2_04_bench_${w}_staged${so}.json').read().strip().splitlines()[-1]);print('$w staged',$so,d['ms_per_step'],d['value'])"; done
done
MVAE_TIMELINE=1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2> gpurun_out/r2_04_timeline.txt
for a in "bar" "bulk 8192" "bulk 90000"; do
  for tool in memcheck synccheck racecheck; do
    timeout 120 compute-sanitizer --tool $tool --print-limit 3 scripts/sanitizer_probe $a > "gpurun_out/r2_04_probe_${tool}_${a// /_}.log" 2>&1
    echo "$tool $a rc=$?"; grep -E "probe|SUMMARY|Invalid|Barrier error|Race" "gpurun_out/r2_04_probe_${tool}_${a// /_}.log" | head -4
  done
done
