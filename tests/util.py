"""Shared helpers for the parity tests: build the same model + batch for the oracle and for the engine."""
from __future__ import annotations

import numpy as np
import torch

from midi_vae_b200 import EngineConfig, initial_weights, synth
from oracle import midivae_oracle as O


def make_cfgs(T=16, H=64, L=16, ne=2, nd=2, feedback="as_wired", gate="hard_sigmoid", variant="recurrentshop_recalled", precision="fp32", max_batch=8,
              rnn_mode="auto", lr=2e-4, cell_type="LSTM"):
    ecfg = EngineConfig(input_length=T, lstm_size=H, latent_rep_size=L, num_layers_encoder=ne, num_layers_decoder=nd, gate_act=gate,
                        dec_cell_variant=variant, decoder_feedback=feedback, precision=precision, max_batch=max_batch, rnn_mode=rnn_mode,
                        learning_rate=lr, cell_type=cell_type)
    ocfg = O.OracleConfig(input_length=T, lstm_size=H, latent_rep_size=L, num_layers_encoder=ne, num_layers_decoder=nd, gate_act=gate,
                          dec_cell_variant=variant, decoder_feedback=feedback, learning_rate=lr, cell_type=cell_type)
    return ecfg, ocfg


def make_weights(ecfg, seed=42, jitter=0.05):
    """Keras-initialised weights plus a small jitter so that zero-initialised biases carry signal."""
    w = initial_weights(ecfg, seed)
    rng = np.random.default_rng(seed + 1)
    return {k: (v + jitter * rng.standard_normal(v.shape)).astype(np.float32) for k, v in w.items()}


def to_torch(w, dtype=torch.float64):
    return {k: torch.tensor(v, dtype=dtype) for k, v in w.items()}


def make_batch(ecfg, n, seed=7, hist_scale=0.1, eps_std=0.01, weights=False):
    r = synth.make_batch(n, ecfg.input_length, seed=seed)
    rng = np.random.default_rng(seed + 100)
    hist = (rng.standard_normal((n, ecfg.latent_rep_size)) * hist_scale).astype(np.float32)
    eps = synth.make_eps(n, ecfg.latent_rep_size, seed, eps_std)
    w = None
    if weights:
        w = np.ones((n, ecfg.input_length), np.float32)
        w[r.pitch == synth.SILENT] = 0.5
        w[0, 1] = 0.0
    return r, hist, eps, w


def oracle_inputs(ocfg, r, hist, eps, w, dtype=torch.float64):
    X, I, V, C = [torch.tensor(a, dtype=dtype) for a in r.dense(np.float64)]
    sw = None if w is None else (torch.tensor(w, dtype=dtype), None, None, None)
    return X, I, V, C, torch.tensor(hist, dtype=dtype), torch.tensor(eps, dtype=dtype), sw


def rel_err(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))
