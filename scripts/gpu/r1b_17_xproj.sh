#!/bin/bash
mkdir -p gpurun_out
echo "=== parity"
timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "persistent_rnn or bf16 or style_transfer" -p no:cacheprovider 2>&1 | grep "assert\|Error\|passed\|failed\|FAILED\|timeout\|trap" | head -20
for fx in 1; do
echo "=== fuse_xproj=$fx bench cfg3"
MVAE_FUSE_XPROJ=$fx MVAE_TIMELINE=1 timeout 600 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r1b_17_bench_$fx.log 2> gpurun_out/r1b_17_bench_$fx.err
tail -1 gpurun_out/r1b_17_bench_$fx.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['class_ms'])"
done
grep timeline gpurun_out/r1b_17_bench_1.err | head -70
