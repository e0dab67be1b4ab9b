#!/bin/bash
# round 2, call 19: K-split-major tile order in the split-K GEMMs (L2 reuse across the tiles of a wave)
set -x
for rep in 1 2 3; do
  for v in 1 2; do
    MVAE_GEMM_RASTER=$v python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('raster',$v,round(d['ms_per_step'],3),round(d['value']),d['roofline']['class_ms'])"
  done
done
python -c "
from midi_vae_b200 import _lib
rc = _lib.load().mvae_selftest_gemm(0, 0); print('selftest rc', rc)" 2>&1 | tail -2
