#!/bin/bash
mkdir -p gpurun_out
echo "=== parity (branch on)"
timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "persistent_rnn or bf16 or style_transfer" -p no:cacheprovider 2>&1 | grep "assert\|Error\|passed\|failed\|FAILED\|timeout\|trap" | head -20
for br in 1 0; do
echo "=== branch=$br bench cfg3"
MVAE_BRANCH=$br timeout 600 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r1b_15_bench_$br.log 2> gpurun_out/r1b_15_bench_$br.err
tail -1 gpurun_out/r1b_15_bench_$br.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['class_ms'])"
tail -3 gpurun_out/r1b_15_bench_$br.err
done
echo "=== branch=1 side_sms=36"
MVAE_SIDE_SMS=36 timeout 600 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['class_ms'])"
echo "=== cfg2"
timeout 600 python bench.py --workload cfg2 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['class_ms'])"
