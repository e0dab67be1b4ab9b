#!/bin/bash
# round 2, call 5: bwd4 stash prefetch one step ahead + stmatrix message staging (ISA probe first), free-running decode as a CUDA graph
set -x
mkdir -p gpurun_out
scripts/isa_probe > gpurun_out/r2_05_isa_probe.log 2>&1; cat gpurun_out/r2_05_isa_probe.log
python -m pytest tests -m gpu -q -s --durations=5 2>&1 | tail -60 > gpurun_out/r2_05_pytest.log
tail -4 gpurun_out/r2_05_pytest.log
MVAE_CLB_STM=1 python -m pytest tests/test_gpu_parity_bench_shapes.py tests/test_gpu_parity.py -m gpu -q -s -k "cfg3 or persistent_rnn or chunked or overlap or bf16" 2>&1 | tail -40 > gpurun_out/r2_05_pytest_stm.log
tail -4 gpurun_out/r2_05_pytest_stm.log
for stm in 0 1; do
  for pf in 2 4; do
    MVAE_CLB_STM=$stm MVAE_CLB_PREFETCH=$pf python bench.py --steps 15 --warmup 4 --no-cpu-baseline --no-e2e > gpurun_out/r2_05_bench_stm${stm}_pf${pf}.json 2> gpurun_out/r2_05_bench_stm${stm}_pf${pf}.err
    python -c "import json;d=json.loads(open('gpurun_out/r2_05_bench_stm${stm}_pf${pf}.json').read().strip().splitlines()[-1]);print('stm',$stm,'pf',$pf,d['ms_per_step'],d['value'],d['roofline']['kernel_ms_per_step'])"
  done
done
for gr in 1 0; do
  MVAE_STEPWISE_GRAPH=$gr python bench.py --workload cfg4 --infer-feedback free_running --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_05_bench_cfg4_free_graph${gr}.json 2> gpurun_out/r2_05_bench_cfg4_free_graph${gr}.err
  python -c "import json;d=json.loads(open('gpurun_out/r2_05_bench_cfg4_free_graph${gr}.json').read().strip().splitlines()[-1]);print('cfg4 free_running graph',$gr,d['ms_per_step'],d['value'],d.get('e2e'))"
done
MVAE_CLB_STM=1 MVAE_TIMELINE=1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2> gpurun_out/r2_05_timeline_stm1.txt
