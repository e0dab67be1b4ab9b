#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider 2>&1 | tail -6
timeout 900 python scripts/sweep_rnn.py sweep 2>&1 | tee gpurun_out/sweep_rnn.jsonl | cut -c1-600
timeout 600 python scripts/sweep_rnn.py infer 2>&1 | tee gpurun_out/infer_cfg4.jsonl | cut -c1-400
