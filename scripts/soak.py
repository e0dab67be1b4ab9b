"""Soak of the cluster kernels: N train-step evaluations of ONE batch from ONE set of weights.

    python scripts/soak.py [iterations=500] [H=512] [T=256] [B=512]

Every iteration restores the weights and the optimiser state, runs forward + backward, and compares with iteration 0:
  * the ten metrics (forward pass only: recurrences, heads, losses) must be BIT-IDENTICAL -- the forward path has no atomics;
  * every gradient tensor must agree to 1e-5 of its max (the weight-gradient GEMMs split K and accumulate with fp32 atomics, so the summation
    order, and nothing else, may differ between runs).
Any hang, NaN or mismatch of the mbarrier / multicast / ACK protocols of rec_cluster_fwd2 / bwd4 shows up here as a differing bit pattern.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from midi_vae_b200 import Engine, METRIC_KEYS  # noqa: E402
from tests import util                          # noqa: E402


def main():
    it = int(sys.argv[1]) if len(sys.argv) > 1 else 500
    H = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    T = int(sys.argv[3]) if len(sys.argv) > 3 else 256
    B = int(sys.argv[4]) if len(sys.argv) > 4 else 512
    ecfg, _ = util.make_cfgs(T=T, H=H, L=256, feedback="teacher_forced", precision="bf16", max_batch=B)
    eng = Engine(ecfg, 0)
    w = util.make_weights(ecfg)
    eng.set_weights(w)
    r, hist, eps, sw = util.make_batch(ecfg, B, weights=True)
    names = ["lstm_1/recurrent_kernel", "lstm_2/kernel", "lstm_meta_velocity/kernel", "notes/cell_1/recurrent_kernel", "notes/cell_2/kernel", "dec_init/notes_l1_s1/kernel"]
    ref_bits, ref_g = None, None
    worst, bad = 0.0, 0
    t0 = time.time()
    for i in range(it):
        eng.set_weights(w); eng.reset_optimizer()
        m = eng.train_on_batch(r.pitch, r.instr, r.velocity, r.style, hist, eps, sw)
        bits = np.array([m[k] for k in METRIC_KEYS], np.float32).view(np.uint32)
        if i % 25 == 0 or i == it - 1:
            g = eng.get_grads()
            if ref_g is None:
                ref_g = {k: g[k].copy() for k in names}
            for k in names:
                worst = max(worst, float(np.abs(g[k] - ref_g[k]).max() / (np.abs(ref_g[k]).max() + 1e-30)))
        if ref_bits is None:
            ref_bits = bits
        elif not np.array_equal(bits, ref_bits):
            bad += 1
            print(f"iteration {i}: metrics differ from iteration 0: {dict(zip(METRIC_KEYS, np.array(bits).view(np.float32)))}")
    print(f"soak H={H} T={T} B={B}: {it} iterations in {time.time() - t0:.1f} s; metric bit patterns differing from iteration 0: {bad}; "
          f"worst gradient deviation {worst:.2e} of the tensor max (fp32 atomics order); loss {m['loss']:.6f}")
    eng.close()
    sys.exit(1 if bad or worst > 1e-5 or not np.isfinite(m["loss"]) else 0)


if __name__ == "__main__":
    main()
