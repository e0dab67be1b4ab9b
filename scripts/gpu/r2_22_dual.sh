#!/bin/bash
# round 2, call 22: dU and dW of a dense-input recurrence in one split-K launch (dG streamed once)
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x --durations=3 2>&1 | tail -6 > gpurun_out/r2_22_pytest.log; tail -3 gpurun_out/r2_22_pytest.log
for rep in 1 2 3; do
  for v in 0 1; do
    MVAE_WGRAD_DUAL=$v python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('dual',$v,round(d['ms_per_step'],3),round(d['value']),d['roofline']['class_ms'])"
  done
done
python bench.py --workload cfg2 --steps 50 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('cfg2',round(d['ms_per_step'],3),round(d['value']))"
