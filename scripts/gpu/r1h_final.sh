#!/bin/bash
# banked deliverables of the DEFAULT build (no environment switches): full GPU test suite, default bench line, reference arm, cfg2,
# timeline, ncu launch list, ncu full captures (weight-gradient / projection GEMMs, forward cluster kernel)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
show='import sys,json; d=json.loads(sys.stdin.read()); print(round(d["value"]), "seq/s", round(d["ms_per_step"],3), "ms", d["roofline"]["class_ms"])'
echo "=== default bench"
timeout 900 python bench.py > gpurun_out/r1h_bench_default.log 2> gpurun_out/r1h_bench_default.err
tail -1 gpurun_out/r1h_bench_default.log | cut -c1-200; tail -1 gpurun_out/r1h_bench_default.log | python -c "$show"
echo "=== reference arm"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | tee gpurun_out/r1h_bench_reference.log | cut -c1-200
echo "=== cfg2"
timeout 300 python bench.py --workload cfg2 --steps 20 --warmup 5 --no-cpu-baseline 2> gpurun_out/r1h_cfg2.err | tail -1 | tee gpurun_out/r1h_bench_cfg2.log | python -c "$show"
echo "=== timeline"
MVAE_TIMELINE=1 timeout 300 python bench.py --workload cfg3 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r1h_timeline.log 2> gpurun_out/r1h_timeline.err
grep timeline gpurun_out/r1h_timeline.err | tail -61 > gpurun_out/r1h_timeline.txt; wc -l gpurun_out/r1h_timeline.txt
echo "=== ncu launch list (2 steps)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/r1h_launches_cfg3.csv python scripts/scripts_one_step.py persistent 2 > gpurun_out/r1h_ncu_list.log 2>&1
tail -1 gpurun_out/r1h_ncu_list.log; wc -l gpurun_out/r1h_launches_cfg3.csv
echo "=== ncu full: GEMMs of the second step"
timeout 900 ncu --set full --clock-control none -k regex:"gemm_tc_kernel" -s 42 -c 16 -o gpurun_out/r1h_prof_gemm python scripts/scripts_one_step.py persistent 2 > gpurun_out/r1h_ncu_gemm.log 2>&1
tail -1 gpurun_out/r1h_ncu_gemm.log; ls -la gpurun_out/r1h_prof_gemm.ncu-rep
echo "=== ncu full: cluster kernels of the second step"
timeout 900 ncu --set full --clock-control none -k regex:"rec_cluster" -s 12 -c 4 -o gpurun_out/r1h_prof_rec python scripts/scripts_one_step.py persistent 2 > gpurun_out/r1h_ncu_rec.log 2>&1
tail -1 gpurun_out/r1h_ncu_rec.log; ls -la gpurun_out/r1h_prof_rec.ncu-rep
for r in gemm rec; do ncu -i gpurun_out/r1h_prof_$r.ncu-rep --page raw --csv > gpurun_out/r1h_prof_${r}_raw.csv 2>/dev/null; done
du -sm gpurun_out; if [ $(du -sm gpurun_out | cut -f1) -gt 55 ]; then rm -f gpurun_out/r1h_prof_gemm.ncu-rep; fi; du -sm gpurun_out
