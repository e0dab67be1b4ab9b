#!/bin/bash
# round-1 measurement set: ncu launch list of two cfg3 steps, full ncu captures of the cluster recurrence kernels, inference + sweep
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
echo "=== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r1c_launches_cfg3.csv python scripts/scripts_one_step.py persistent 2 > gpurun_out/r1c_ncu_list.log 2>&1
tail -1 gpurun_out/r1c_ncu_list.log; wc -l gpurun_out/r1c_launches_cfg3.csv
echo "=== ncu full: cluster recurrence kernels"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"rec_cluster" -s 16 -c 6 -o gpurun_out/r1c_prof_rec python scripts/scripts_one_step.py persistent 2 > gpurun_out/r1c_ncu_rec.log 2>&1
tail -2 gpurun_out/r1c_ncu_rec.log; ls -la gpurun_out/r1c_prof_rec.ncu-rep
echo "=== inference"
timeout 900 python scripts/sweep_rnn.py infer 2>&1 | tail -6 | tee gpurun_out/r1c_infer.jsonl
echo "=== sweep"
timeout 1200 python scripts/sweep_rnn.py sweep 2>&1 | tail -6 | tee gpurun_out/r1c_sweep.jsonl
