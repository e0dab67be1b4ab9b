#!/bin/bash
# round 2, call 15: column sums with few long-running blocks (atomic contention fix): suite, bench, timeline
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x --durations=3 2>&1 | tail -8 > gpurun_out/r2_15_pytest.log; tail -3 gpurun_out/r2_15_pytest.log
for i in 1 2 3; do python bench.py --steps 15 --warmup 4 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('cfg3',round(d['ms_per_step'],3),round(d['value']),d['roofline']['class_ms'])"; done
python bench.py --workload cfg2 --steps 50 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('cfg2',round(d['ms_per_step'],3),round(d['value']))"
MVAE_TIMELINE=2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2> gpurun_out/r2_15_timeline.txt
awk '$4>4.5 && $2=="gemm"' gpurun_out/r2_15_timeline.txt | sort -k4 -n | head -60
grep "adam\|rec_bwd" gpurun_out/r2_15_timeline.txt | tail -8
