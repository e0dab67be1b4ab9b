import sys, os
sys.path.insert(0, '.')
import numpy as np
from midi_vae_b200 import Engine, EngineConfig, initial_weights, synth
wl = dict(T=256, H=512, L=256, B=512)
mode = sys.argv[1] if len(sys.argv) > 1 else "persistent"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cfg = EngineConfig(input_length=wl["T"], lstm_size=wl["H"], latent_rep_size=wl["L"], decoder_feedback="teacher_forced", precision="bf16", rnn_mode=mode, max_batch=wl["B"])
eng = Engine(cfg, 0); eng.set_weights(initial_weights(cfg, 42))
r = synth.make_batch(wl["B"], wl["T"], seed=1); eps = synth.make_eps(wl["B"], wl["L"], 1)
for i in range(steps):
    m = eng.train_on_batch(r.pitch, r.instr, r.velocity, r.style, None, eps)
print(m["loss"])
