#!/bin/bash
mkdir -p gpurun_out
for ns in 0 1; do
echo "=== parity nswap=$ns"
MVAE_CL_NSWAP=$ns timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "persistent_rnn and (shape1 or shape5 or shape6)" -p no:cacheprovider 2>&1 | grep -v "^$" | grep "assert\|Error\|passed\|failed\|FAILED\|timeout\|trap" | head -40
done
