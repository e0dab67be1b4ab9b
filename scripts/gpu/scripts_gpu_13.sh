#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider -k "persistent_rnn and (shape3 or shape4)" 2>&1 | grep -E "Error|assert|FAILED|passed|failed" | cut -c1-300 | head -40
