#!/bin/bash
# round 2, call 3: the two re-designed cfg4 parity tests, sanitizers (T = 8), and the branch-wave experiment (MVAE_BRANCH_BWD_NCL) at cfg3
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity_bench_shapes.py -m gpu -q -s -k "cfg4" 2>&1 | tail -30 > gpurun_out/r2_03_pytest_cfg4.log
tail -5 gpurun_out/r2_03_pytest_cfg4.log
for v in 0 2 1 3; do
  for sms in 0 100; do
    MVAE_BRANCH_BWD_NCL=$v MVAE_SIDE_SMS=$sms python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_03_bench_ncl${v}_sms${sms}.json 2> gpurun_out/r2_03_bench_ncl${v}_sms${sms}.err
    python -c "import json;d=json.loads(open('gpurun_out/r2_03_bench_ncl${v}_sms${sms}.json').read().strip().splitlines()[-1]);print('ncl',$v,'sms',$sms,d['ms_per_step'],d['value'])"
  done
done
MVAE_BRANCH_BWD_NCL=2 MVAE_TIMELINE=1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2> gpurun_out/r2_03_timeline_ncl2.txt
python scripts/sanitize_case.py 512 8 72 > gpurun_out/r2_03_plain_512.log 2>&1; tail -1 gpurun_out/r2_03_plain_512.log
for tool in memcheck synccheck racecheck; do
  for H in 512 256; do
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_case.py $H 8 72 > gpurun_out/r2_03_san_${tool}_${H}.log 2>&1
    echo "$tool $H rc=$?"; tail -4 gpurun_out/r2_03_san_${tool}_${H}.log
  done
done
