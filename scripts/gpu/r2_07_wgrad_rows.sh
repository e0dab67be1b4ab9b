#!/bin/bash
# round 2, call 7: single-pass weight gradients (one-hot / scalar inputs), self-addressed messages copied locally, explicit convergence before named
# barriers: full suite, A/B bench, sanitizers again, soak
set -x
mkdir -p gpurun_out
rm -f gpurun_out/parity_bench_shapes.jsonl
python -m pytest tests -m gpu -q -s --durations=5 2>&1 | tail -60 > gpurun_out/r2_07_pytest.log
tail -4 gpurun_out/r2_07_pytest.log
MVAE_CLB_STM=1 python -m pytest tests/test_gpu_parity_bench_shapes.py tests/test_gpu_parity.py -m gpu -q -k "cfg3 or persistent_rnn or chunked or overlap or bf16" 2>&1 | tail -5 > gpurun_out/r2_07_pytest_stm.log
tail -2 gpurun_out/r2_07_pytest_stm.log
for cfgs in "0 1" "1 1" "1 0"; do
  set -- $cfgs
  MVAE_CLB_STM=$1 MVAE_WGRAD_ROWS=$2 python bench.py --steps 15 --warmup 4 --no-cpu-baseline --no-e2e > gpurun_out/r2_07_bench_stm$1_rows$2.json 2> gpurun_out/r2_07_bench_stm$1_rows$2.err
  python -c "import json;d=json.loads(open('gpurun_out/r2_07_bench_stm$1_rows$2.json').read().strip().splitlines()[-1]);print('stm',$1,'rows',$2,d['ms_per_step'],d['value'],d['roofline']['kernel_ms_per_step'])"
done
MVAE_CLB_STM=1 python bench.py --workload cfg2 --steps 50 --no-cpu-baseline --no-e2e > gpurun_out/r2_07_bench_cfg2.json 2> gpurun_out/r2_07_bench_cfg2.err; python -c "import json;d=json.loads(open('gpurun_out/r2_07_bench_cfg2.json').read().strip().splitlines()[-1]);print('cfg2',d['ms_per_step'],d['value'])"
MVAE_CLB_STM=1 MVAE_TIMELINE=2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2> gpurun_out/r2_07_timeline_detail.txt
MVAE_CLB_STM=1 python scripts/soak.py 300 > gpurun_out/r2_07_soak_cfg3.log 2>&1; tail -2 gpurun_out/r2_07_soak_cfg3.log
python scripts/soak.py 300 256 64 128 > gpurun_out/r2_07_soak_cfg2.log 2>&1; tail -2 gpurun_out/r2_07_soak_cfg2.log
for tool in memcheck synccheck racecheck; do
  for H in 512 256; do
    MVAE_CLB_STM=1 timeout 900 compute-sanitizer --tool $tool --print-limit 8 python scripts/sanitize_case.py $H 8 72 > gpurun_out/r2_07_san_${tool}_${H}.log 2>&1
    echo "$tool $H rc=$?"; grep -E "SUMMARY|sanitize_case H" gpurun_out/r2_07_san_${tool}_${H}.log | tail -3
  done
done
