#!/bin/bash
# banked deliverables: full GPU test suite, default bench line, ncu full capture of the backward cluster kernel + the GEMM kernel
mkdir -p gpurun_out
echo "=== full gpu tests"
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -4
echo "=== default bench"
timeout 900 python bench.py > gpurun_out/r1c_bench_default.log 2> gpurun_out/r1c_bench_default.err
tail -1 gpurun_out/r1c_bench_default.log | cut -c1-400
echo "=== reference arm"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
echo "=== ncu full: backward cluster kernel"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"rec_cluster_bwd" -s 8 -c 4 -o gpurun_out/r1c_prof_bwd python scripts/scripts_one_step.py persistent 2 > gpurun_out/r1c_ncu_bwd.log 2>&1
tail -1 gpurun_out/r1c_ncu_bwd.log; ls -la gpurun_out/r1c_prof_bwd.ncu-rep
echo "=== ncu full: gemm"
timeout 1500 ncu --set full --clock-control none -k regex:"gemm_tc_kernel" -s 40 -c 12 -o gpurun_out/r1c_prof_gemm python scripts/scripts_one_step.py persistent 2 > gpurun_out/r1c_ncu_gemm.log 2>&1
tail -1 gpurun_out/r1c_ncu_gemm.log; ls -la gpurun_out/r1c_prof_gemm.ncu-rep
