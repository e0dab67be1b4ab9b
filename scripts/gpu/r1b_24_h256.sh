#!/bin/bash
python - <<'PY'
import sys, json, os
sys.path.insert(0, '.'); sys.path.insert(0, 'scripts')
import sweep_rnn
for mx in (8, 0, 16):
    os.environ["MVAE_CL_NG1_MAX"] = str(mx)
    for T, H, L, B in ((64, 256, 100, 128), (256, 256, 256, 512), (64, 256, 100, 1024)):
        ms, k = sweep_rnn.train_ms(T, H, L, B, "persistent", steps=5)
        print(mx, (T, H, L, B), round(ms, 3), round(B / ms * 1e3), k, flush=True)
PY
