#!/bin/bash
# round 2, call 27: full GPU suite, smoke and every bench line of the build with the CTA-pair weight-gradient GEMM / vector reductions / unrolled column sums
set -x
mkdir -p gpurun_out
rm -f gpurun_out/parity_bench_shapes.jsonl
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader; nproc
timeout 900 python -m pytest tests -m gpu -q -s --durations=8 2>&1 | tail -70 > gpurun_out/r2_27_pytest.log; tail -4 gpurun_out/r2_27_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_27_smoke.log 2>&1; tail -5 gpurun_out/r2_27_smoke.log
timeout 600 python bench.py > gpurun_out/r2_27_bench_cfg3.json 2> gpurun_out/r2_27_bench.err; tail -c 300 gpurun_out/r2_27_bench_cfg3.json
timeout 300 python bench.py --workload cfg2 --steps 50 > gpurun_out/r2_27_bench_cfg2.json 2>> gpurun_out/r2_27_bench.err
timeout 300 python bench.py --workload cfg4 --steps 20 > gpurun_out/r2_27_bench_cfg4.json 2>> gpurun_out/r2_27_bench.err
timeout 300 python bench.py --workload cfg4 --cfg4-shape cfg2 --steps 20 > gpurun_out/r2_27_bench_cfg4_cfg2shape.json 2>> gpurun_out/r2_27_bench.err
timeout 300 python bench.py --workload cfg4 --infer-feedback free_running --steps 10 --warmup 3 > gpurun_out/r2_27_bench_cfg4_free_running.json 2>> gpurun_out/r2_27_bench.err
timeout 300 python bench.py --workload refdefault --steps 10 --warmup 3 > gpurun_out/r2_27_bench_refdefault_gru.json 2>> gpurun_out/r2_27_bench.err
timeout 300 python bench.py --workload cfg5 --steps 3 --warmup 3 > gpurun_out/r2_27_bench_cfg5.json 2>> gpurun_out/r2_27_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_27_bench_reference.json 2>> gpurun_out/r2_27_bench.err
MVAE_TIMELINE=2 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2> gpurun_out/r2_27_timeline.txt
for f in cfg3 cfg2 cfg4 cfg4_cfg2shape cfg4_free_running refdefault_gru cfg5 reference; do python -c "
import json
d=json.loads(open('gpurun_out/r2_27_bench_$f.json').read().strip().splitlines()[-1])
print('$f', round(d['ms_per_step'],3), round(d['value'],1), (d.get('e2e') or {}).get('value'), (d.get('roofline') or {}).get('frac'), ((d.get('roofline') or {}).get('step') or {}).get('frac'))"; done
for rep in 1 2; do timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('repeat',round(d['ms_per_step'],3),round(d['value']),d['clocks'])"; done
