#!/bin/bash
# round 2, call 14: style classifiers (incl. the shipped checkpoints), history from the batch itself, whole suite again, head weight-gradient timeline
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_classifier.py -m gpu -q -s 2>&1 | tail -40 > gpurun_out/r2_14_pytest_cls.log; tail -6 gpurun_out/r2_14_pytest_cls.log
python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "history_from_the_batch or packed" 2>&1 | tail -30 > gpurun_out/r2_14_pytest_hist.log; tail -4 gpurun_out/r2_14_pytest_hist.log
python -m pytest tests -m gpu -q --durations=5 2>&1 | tail -30 > gpurun_out/r2_14_pytest.log; tail -4 gpurun_out/r2_14_pytest.log
MVAE_TIMELINE=2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2> gpurun_out/r2_14_timeline.txt
grep heads gpurun_out/r2_14_timeline.txt
python bench.py --steps 15 --warmup 4 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('cfg3',d['ms_per_step'],d['value'])"
