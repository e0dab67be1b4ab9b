#!/bin/bash
# round 2, call 18: TIMING experiment -- decoder weight gradients deferred into the next step's encoder forward (gradients are not valid in this mode)
set -x
for rep in 1 2; do
  for v in 0 1; do
    MVAE_EXP_DEFER=$v python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('defer',$v,round(d['ms_per_step'],3),round(d['value']))"
  done
done
MVAE_EXP_DEFER=1 MVAE_TIMELINE=1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2> gpurun_out/r2_18_timeline_defer.txt
grep -c timeline gpurun_out/r2_18_timeline_defer.txt; grep "rec_fwd\|rec_bwd\|adam" gpurun_out/r2_18_timeline_defer.txt | awk '$6>0.3' 
