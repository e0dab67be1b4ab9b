#!/bin/bash
# round 2, call 16: weight-gradient stream one priority level above the branch recurrences (a pending branch cluster blocks same-priority grids behind it?)
set -x
mkdir -p gpurun_out
python -c "import torch; print(torch.cuda.get_device_properties(0).name)"; 
for rep in 1 2 3; do
  for v in 0 1; do
    MVAE_SIDE_PRIO=$v python bench.py --steps 15 --warmup 4 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('side_prio',$v,round(d['ms_per_step'],3),round(d['value']))"
  done
done
MVAE_SIDE_PRIO=1 MVAE_TIMELINE=2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2> gpurun_out/r2_16_timeline_prio1.txt
grep "wgrad\|rec_bwd\|adam" gpurun_out/r2_16_timeline_prio1.txt | sort -k4 -n | tail -45
