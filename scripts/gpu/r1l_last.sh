#!/bin/bash
# last call of the round: smoke() + the full GPU suite on the final build
mkdir -p gpurun_out
echo "=== full gpu tests"
timeout 600 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -25
