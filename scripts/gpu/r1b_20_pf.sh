#!/bin/bash
for pf in 0 2 4 8; do
echo "=== bwd l2 prefetch distance $pf"
MVAE_CLB_PREFETCH=$pf timeout 600 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['class_ms'])"
done
