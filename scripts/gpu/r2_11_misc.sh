#!/bin/bash
# round 2, call 11: last look at the off-by-default pipelines before they are deleted, GRU reference-default workload with / without the step graph,
# staged epilogue on / off at cfg4
set -x
mkdir -p gpurun_out
show() { python -c "import json,sys;d=json.loads(open('$1').read().strip().splitlines()[-1]);print('$2',d['ms_per_step'],d['value'])"; }
for v in "MVAE_XW_OVERLAP=4" "MVAE_XW_OVERLAP=8" "MVAE_CHUNKS=4 MVAE_CHUNKS_BWD=4" "MVAE_CHUNKS=2" "MVAE_CLB_NG=1" "MVAE_BRANCH_BWD_NCL=3"; do
  f=gpurun_out/r2_11_bench_$(echo $v | tr ' =' '__').json
  env $v python bench.py --steps 15 --warmup 4 --no-cpu-baseline --no-e2e > $f 2>/dev/null; show $f "$v"
done
python bench.py --steps 15 --warmup 4 --no-cpu-baseline --no-e2e > gpurun_out/r2_11_bench_default.json 2>/dev/null; show gpurun_out/r2_11_bench_default.json default
for g in 1 0; do
  MVAE_STEP_GRAPH=$g python bench.py --workload refdefault --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_11_bench_refdefault_graph$g.json 2> gpurun_out/r2_11_bench_refdefault_graph$g.err; show gpurun_out/r2_11_bench_refdefault_graph$g.json "refdefault graph $g"
  MVAE_STEP_GRAPH=$g python bench.py --workload cfg1 --precision fp32 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_11_bench_cfg1_fp32_graph$g.json 2> gpurun_out/r2_11_bench_cfg1_fp32_graph$g.err; show gpurun_out/r2_11_bench_cfg1_fp32_graph$g.json "cfg1 fp32 graph $g"
done
for so in 1 0 1 0; do
  MVAE_GEMM_STAGED_OUT=$so python bench.py --workload cfg4 --steps 20 --no-cpu-baseline --no-e2e > gpurun_out/r2_11_bench_cfg4_staged$so.json 2>/dev/null; show gpurun_out/r2_11_bench_cfg4_staged$so.json "cfg4 staged $so"
done
python -m pytest tests -m gpu -q -x -k "gru or trajectory or fp32" 2>&1 | tail -4
