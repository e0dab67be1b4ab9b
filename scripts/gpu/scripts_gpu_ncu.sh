#!/bin/bash
# ncu: launch list of one cfg3 step (persistent) + full captures of the dominant kernels
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_cfg3.csv python scripts_one_step.py persistent 2 > gpurun_out/ncu_list.log 2>&1
tail -1 gpurun_out/ncu_list.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"rec_persist_kernel|rec_bwd_ksplit" -s 8 -c 6 -o gpurun_out/prof_rec python scripts_one_step.py persistent 2 > gpurun_out/ncu_rec.log 2>&1
tail -1 gpurun_out/ncu_rec.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 100 -c 40 -o gpurun_out/prof_gemm python scripts_one_step.py persistent 2 > gpurun_out/ncu_gemm.log 2>&1
tail -1 gpurun_out/ncu_gemm.log
