#!/bin/bash
# round 2, call 9: 256 x 256 weight-gradient tiles (self test first), fused scalar-input weight gradients, stmatrix default; synccheck probe with a warp parked at __syncthreads
set -x
mkdir -p gpurun_out
python -c "
from midi_vae_b200 import _lib
import sys
rc = _lib.load().mvae_selftest_gemm(0, 0); print('selftest rc', rc); sys.exit(rc)" > gpurun_out/r2_09_selftest.log 2>&1; tail -3 gpurun_out/r2_09_selftest.log
python -m pytest tests -m gpu -q -x --durations=5 2>&1 | tail -30 > gpurun_out/r2_09_pytest.log
tail -4 gpurun_out/r2_09_pytest.log
for bm in 1 0; do
  MVAE_GEMM_BM256=$bm python bench.py --steps 15 --warmup 4 --no-cpu-baseline --no-e2e > gpurun_out/r2_09_bench_bm$bm.json 2> gpurun_out/r2_09_bench_bm$bm.err
  python -c "import json;d=json.loads(open('gpurun_out/r2_09_bench_bm$bm.json').read().strip().splitlines()[-1]);print('bm256',$bm,d['ms_per_step'],d['value'],d['roofline']['kernel_ms_per_step'])"
done
MVAE_TIMELINE=2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2> gpurun_out/r2_09_timeline_detail.txt
MVAE_SIDE_SMS=148 python bench.py --steps 15 --warmup 4 --no-cpu-baseline --no-e2e > gpurun_out/r2_09_bench_side148.json 2>/dev/null; python -c "import json;d=json.loads(open('gpurun_out/r2_09_bench_side148.json').read().strip().splitlines()[-1]);print('side148',d['ms_per_step'],d['value'])"
MVAE_SIDE_SMS=60 python bench.py --steps 15 --warmup 4 --no-cpu-baseline --no-e2e > gpurun_out/r2_09_bench_side60.json 2>/dev/null; python -c "import json;d=json.loads(open('gpurun_out/r2_09_bench_side60.json').read().strip().splitlines()[-1]);print('side60',d['ms_per_step'],d['value'])"
python bench.py --workload cfg2 --steps 50 --no-cpu-baseline --no-e2e > gpurun_out/r2_09_bench_cfg2.json 2>/dev/null; python -c "import json;d=json.loads(open('gpurun_out/r2_09_bench_cfg2.json').read().strip().splitlines()[-1]);print('cfg2',d['ms_per_step'],d['value'])"
timeout 120 compute-sanitizer --tool synccheck --print-limit 3 scripts/sanitizer_probe barwait > gpurun_out/r2_09_probe_synccheck_barwait.log 2>&1; grep -E "probe|SUMMARY|Barrier" gpurun_out/r2_09_probe_synccheck_barwait.log | head -4
