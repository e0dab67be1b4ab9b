"""Generate the committed golden vectors from the fp64 CPU oracle  (oracle-derived, NOT reference-derived:
the reference ships no fixtures and cannot run here).

    python tests/golden/make_golden.py

Writes tests/golden/cfg1_step.npz: one train step + one style-transfer call at BASELINE config[0] shapes
(seq_len=16, hidden=64, latent=16, batch=8) with teacher-forced decoders: packed inputs, weights, the ten
metrics, every gradient tensor, argmax outputs and their top-2 margins.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from midi_vae_b200 import METRIC_KEYS  # noqa: E402
from oracle import midivae_oracle as O  # noqa: E402
from tests import util  # noqa: E402


def main():
    feedback, variant = "teacher_forced", "recurrentshop_recalled"      # the reference decoder cell (recurrentshop LSTMCell as recalled)
    ecfg, ocfg = util.make_cfgs(T=16, H=64, L=16, feedback=feedback, variant=variant, max_batch=8)
    w = util.make_weights(ecfg, seed=42, jitter=0.1)
    r, hist, eps, _ = util.make_batch(ecfg, 8, seed=1235)
    p = util.to_torch(w)
    X, I, V, C, th, te, _ = util.oracle_inputs(ocfg, r, hist, eps, None)
    m, g, _ = O.loss_and_grads(ocfg, p, X, I, V, C, th, te)
    st = O.style_transfer(ocfg, p, X, I, V, 0, 1, None, "as_wired")
    out = dict(feedback=np.array(feedback), variant=np.array(variant), pitch=r.pitch, instr=r.instr, velocity=r.velocity, style=r.style, hist=hist, eps=eps,
               metrics=np.array([m[k] for k in METRIC_KEYS]), st_pitch=st["pitch"].numpy().astype(np.uint8),
               st_instr=st["instr"].numpy().astype(np.uint8), st_margin=O.top2_margin(st["Yh"]).numpy(),
               st_vel=st["Vh"].numpy()[..., 0].astype(np.float32))
    for k, v in w.items():
        out["w/" + k] = v
    for k, v in g.items():
        out["g/" + k] = v.numpy().astype(np.float32)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cfg1_step.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes; loss", m["loss"])


if __name__ == "__main__":
    main()
