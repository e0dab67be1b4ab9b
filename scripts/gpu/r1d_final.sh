#!/bin/bash
# banked deliverables of the default build: full GPU test suite, default bench line, reference arm, ncu launch list + full capture
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
echo "=== full gpu tests (default build)"
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4
echo "=== default bench"
timeout 900 python bench.py > gpurun_out/r1e_bench_default.log 2> gpurun_out/r1e_bench_default.err
tail -1 gpurun_out/r1e_bench_default.log | cut -c1-300
echo "=== ncu launch list (default build, 2 steps)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/r1e_launches_cfg3.csv python scripts/scripts_one_step.py persistent 2 > gpurun_out/r1e_ncu_list.log 2>&1
tail -1 gpurun_out/r1e_ncu_list.log; wc -l gpurun_out/r1e_launches_cfg3.csv
echo "=== ncu full: backward cluster kernel (quad form)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"rec_cluster_bwd4" -s 4 -c 3 -o gpurun_out/r1e_prof_bwd4 python scripts/scripts_one_step.py persistent 2 > gpurun_out/r1e_ncu_bwd4.log 2>&1
tail -1 gpurun_out/r1e_ncu_bwd4.log; ls -la gpurun_out/r1e_prof_bwd4.ncu-rep
