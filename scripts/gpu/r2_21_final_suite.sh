#!/bin/bash
# round 2, call 21: the cleaned-up build -- full GPU suite, smoke, every bench line, sanitizers, soak
set -x
mkdir -p gpurun_out
rm -f gpurun_out/parity_bench_shapes.jsonl
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader; nproc
python -m pytest tests -m gpu -q -s --durations=8 2>&1 | tail -70 > gpurun_out/r2_21_pytest.log; tail -4 gpurun_out/r2_21_pytest.log
python __graft_entry__.py smoke > gpurun_out/r2_21_smoke.log 2>&1; tail -5 gpurun_out/r2_21_smoke.log
python bench.py > gpurun_out/r2_21_bench_cfg3.json 2> gpurun_out/r2_21_bench.err; tail -c 300 gpurun_out/r2_21_bench_cfg3.json
python bench.py --workload cfg2 --steps 50 > gpurun_out/r2_21_bench_cfg2.json 2>> gpurun_out/r2_21_bench.err
python bench.py --workload cfg4 --steps 20 > gpurun_out/r2_21_bench_cfg4.json 2>> gpurun_out/r2_21_bench.err
python bench.py --workload cfg4 --cfg4-shape cfg2 --steps 20 > gpurun_out/r2_21_bench_cfg4_cfg2shape.json 2>> gpurun_out/r2_21_bench.err
python bench.py --workload cfg4 --infer-feedback free_running --steps 10 --warmup 3 > gpurun_out/r2_21_bench_cfg4_free_running.json 2>> gpurun_out/r2_21_bench.err
python bench.py --workload refdefault --steps 10 --warmup 3 > gpurun_out/r2_21_bench_refdefault_gru.json 2>> gpurun_out/r2_21_bench.err

python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_21_bench_reference.json 2>> gpurun_out/r2_21_bench.err
MVAE_TIMELINE=2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2> gpurun_out/r2_21_timeline.txt
for f in cfg3 cfg2 cfg4 cfg4_cfg2shape cfg4_free_running refdefault_gru reference; do python -c "
import json
d=json.loads(open('gpurun_out/r2_21_bench_$f.json').read().strip().splitlines()[-1])
print('$f', round(d['ms_per_step'],3), round(d['value'],1), (d.get('e2e') or {}).get('value'), (d.get('roofline') or {}).get('frac'), ((d.get('roofline') or {}).get('step') or {}).get('frac'))"; done
python scripts/soak.py 500 > gpurun_out/r2_21_soak_cfg3.log 2>&1; tail -1 gpurun_out/r2_21_soak_cfg3.log
python scripts/soak.py 500 256 64 128 > gpurun_out/r2_21_soak_cfg2.log 2>&1; tail -1 gpurun_out/r2_21_soak_cfg2.log
for tool in memcheck synccheck racecheck; do
  for H in 512 256; do
    timeout 900 compute-sanitizer --tool $tool --print-limit 6 python scripts/sanitize_case.py $H 8 72 > gpurun_out/r2_21_san_${tool}_${H}.log 2>&1
    echo "$tool $H rc=$?"; grep -E "SUMMARY|sanitize_case H" gpurun_out/r2_21_san_${tool}_${H}.log | tail -3
  done
done
for sms in 64 100 116; do
  MVAE_SIDE_SMS=$sms python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('side_sms',$sms,round(d['ms_per_step'],3),round(d['value']))"
done
