#!/bin/bash
# round 2, call 17: side-stream priority x side grid size, 30-step runs
set -x
for rep in 1 2; do
  for v in "0 0" "1 0" "1 40" "1 56" "0 56"; do
    set -- $v
    MVAE_SIDE_PRIO=$1 MVAE_SIDE_SMS=$2 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('side_prio',$1,'side_sms',$2,round(d['ms_per_step'],3),round(d['value']))"
  done
done
