#!/bin/bash
# round 2, call 30: ncu evidence of the final build -- launch list of two cfg3 steps, full-set captures of the cluster recurrences and of the GEMMs of
# the second step, launch list + full set of the GRU cluster kernels at the reference-default shape; full bench line of that shape
set -x
mkdir -p gpurun_out
timeout 600 python bench.py --workload refdefault --steps 20 > gpurun_out/r2_30_bench_refdefault_gru.json 2> gpurun_out/r2_30_bench.err; tail -c 400 gpurun_out/r2_30_bench_refdefault_gru.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/r2_30_launches_cfg3.csv python scripts/one_step.py persistent 2 > gpurun_out/r2_30_ncu_list.log 2>&1
tail -1 gpurun_out/r2_30_ncu_list.log; wc -l gpurun_out/r2_30_launches_cfg3.csv
NG=$(grep -c "gemm_tc_kernel" gpurun_out/r2_30_launches_cfg3.csv); NG=$((NG / 2)); echo "gemm launches per step: $NG"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"rec_cluster" -s 12 -c 12 -o gpurun_out/r2_30_prof_rec python scripts/one_step.py persistent 2 > gpurun_out/r2_30_ncu_rec.log 2>&1
tail -1 gpurun_out/r2_30_ncu_rec.log
timeout 900 ncu --set full --clock-control none -k regex:"gemm_tc_kernel" -s $NG -c $NG -o gpurun_out/r2_30_prof_gemm python scripts/one_step.py persistent 2 > gpurun_out/r2_30_ncu_gemm.log 2>&1
tail -1 gpurun_out/r2_30_ncu_gemm.log
cat > /tmp/gru_step.py <<'PY'
import sys
sys.path.insert(0, '.')
from midi_vae_b200 import Engine, EngineConfig, initial_weights, synth
cfg = EngineConfig(input_length=64, lstm_size=256, latent_rep_size=256, decoder_feedback="as_wired", precision="bf16", max_batch=256, cell_type="GRU")
eng = Engine(cfg, 0); eng.set_weights(initial_weights(cfg, 42))
r = synth.make_batch(256, 64, seed=1); eps = synth.make_eps(256, 256, 1)
for i in range(2):
    m = eng.train_on_batch(r.pitch, r.instr, r.velocity, r.style, None, eps)
print(m["loss"])
PY
MVAE_STEP_GRAPH=0 timeout 600 ncu --set full --import-source on --clock-control none -k regex:"gru_cluster" -s 16 -c 16 -o gpurun_out/r2_30_prof_gru python /tmp/gru_step.py > gpurun_out/r2_30_ncu_gru.log 2>&1
tail -1 gpurun_out/r2_30_ncu_gru.log
for r in rec gemm gru; do ncu -i gpurun_out/r2_30_prof_$r.ncu-rep --page raw --csv > gpurun_out/r2_30_prof_${r}_raw.csv 2>/dev/null; done
ncu -i gpurun_out/r2_30_prof_rec.ncu-rep --page details --csv 2>/dev/null | grep -i -E "stall|Warp Cycles|Issue|Eligible|Bank" | head -200 > gpurun_out/r2_30_prof_rec_details.csv
ncu -i gpurun_out/r2_30_prof_gru.ncu-rep --page details --csv 2>/dev/null | grep -i -E "stall|Warp Cycles|Issue|Eligible|Bank" | head -200 > gpurun_out/r2_30_prof_gru_details.csv
ls -la gpurun_out/r2_30_prof_*; du -sm gpurun_out
rm -f gpurun_out/r2_30_prof_gemm.ncu-rep
if [ $(du -sm gpurun_out | cut -f1) -gt 55 ]; then rm -f gpurun_out/r2_30_prof_rec.ncu-rep; fi
