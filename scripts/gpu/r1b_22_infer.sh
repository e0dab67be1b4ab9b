#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "style_transfer or persistent_rnn" -p no:cacheprovider 2>&1 | grep "assert\|Error\|passed\|failed\|FAILED\|timeout\|trap" | head
timeout 900 python scripts/sweep_rnn.py infer 2>&1 | tail -4 | tee gpurun_out/r1c_infer2.jsonl
