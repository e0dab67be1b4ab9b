#!/bin/bash
# round 2, call 39: the final tree once more: full GPU suite + smoke
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --durations=3 2>&1 | tail -8 > gpurun_out/r2_39_pytest.log; tail -4 gpurun_out/r2_39_pytest.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/r2_39_smoke.log 2>&1; tail -6 gpurun_out/r2_39_smoke.log
