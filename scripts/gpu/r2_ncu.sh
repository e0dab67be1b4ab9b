#!/bin/bash
# round 2: ncu evidence of the final build -- launch list of two cfg3 steps, full-set captures of the quad backward / forward cluster kernels and of
# the GEMMs of the second step, SASS excerpt; summaries go to profiles/r2/
set -x
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/r2_launches_cfg3.csv python scripts/one_step.py persistent 2 > gpurun_out/r2_ncu_list.log 2>&1
tail -1 gpurun_out/r2_ncu_list.log; wc -l gpurun_out/r2_launches_cfg3.csv
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"rec_cluster" -s 12 -c 12 -o gpurun_out/r2_prof_rec python scripts/one_step.py persistent 2 > gpurun_out/r2_ncu_rec.log 2>&1
tail -1 gpurun_out/r2_ncu_rec.log
timeout 900 ncu --set full --clock-control none -k regex:"gemm_tc_kernel" -s 42 -c 40 -o gpurun_out/r2_prof_gemm python scripts/one_step.py persistent 2 > gpurun_out/r2_ncu_gemm.log 2>&1
tail -1 gpurun_out/r2_ncu_gemm.log
for r in rec gemm; do ncu -i gpurun_out/r2_prof_$r.ncu-rep --page raw --csv > gpurun_out/r2_prof_${r}_raw.csv 2>/dev/null; done
ncu -i gpurun_out/r2_prof_rec.ncu-rep --page details --csv 2>/dev/null | grep -i -E "stall|Warp Cycles|Issue|Eligible|Bank" | head -200 > gpurun_out/r2_prof_rec_details.csv
ls -la gpurun_out/r2_prof_*; du -sm gpurun_out
if [ $(du -sm gpurun_out | cut -f1) -gt 50 ]; then rm -f gpurun_out/r2_prof_gemm.ncu-rep; fi
