#!/bin/bash
# round 2, call 38: full bench line (value, e2e, cpu_baseline, roofline) of the reference-default GRU workload on the final build
set -x
mkdir -p gpurun_out
timeout 300 python bench.py --workload refdefault --steps 30 > gpurun_out/r2_38_bench_refdefault_gru.json 2> gpurun_out/r2_38_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/r2_38_bench_refdefault_gru.json').read().strip().splitlines()[-1])
print('gru', round(d['ms_per_step'],3), round(d['value']), round(d['e2e']['value']), d['roofline']['step']['frac'], d['cpu_baseline']['value'], d['clocks'])"
