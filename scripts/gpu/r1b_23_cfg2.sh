#!/bin/bash
for cfg in "0 0" "1 1" "2 1" "1 2"; do
set -- $cfg
echo "=== cfg2 ng fwd=$1 bwd=$2"
MVAE_CL_NG=$1 MVAE_CLB_NG=$2 timeout 600 python bench.py --workload cfg2 --steps 30 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['class_ms'])"
done
MVAE_TIMELINE=1 timeout 600 python bench.py --workload cfg2 --steps 30 --warmup 5 --no-cpu-baseline --no-e2e 2>&1 | grep timeline | awk '{printf "%s %s %s | ", $2, $4, $7} NR%5==0{print ""}'
