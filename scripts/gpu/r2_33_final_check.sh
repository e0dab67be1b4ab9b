#!/bin/bash
# round 2, call 33: the committed tree once more: full GPU suite, smoke (with the GRU cluster case), default bench line
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --durations=5 2>&1 | tail -12 > gpurun_out/r2_33_pytest.log; tail -5 gpurun_out/r2_33_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_33_smoke.log 2>&1; tail -6 gpurun_out/r2_33_smoke.log
timeout 600 python bench.py > gpurun_out/r2_33_bench_cfg3.json 2> gpurun_out/r2_33_bench.err; tail -c 200 gpurun_out/r2_33_bench_cfg3.json
python -c "
import json
d=json.loads(open('gpurun_out/r2_33_bench_cfg3.json').read().strip().splitlines()[-1])
print('cfg3', round(d['ms_per_step'],3), round(d['value'],1), d['e2e']['value'], d['roofline']['frac'], d['roofline']['step']['frac'], d['clocks'])"
