"""Batch marshalling with the reference's positional layouts (vae_definition.py:770-808, 816-865, 880-1045).

The reference builds these lists from module-level globals (``from settings import *``); here the few globals that
matter on the hot path are explicit arguments.  X (N,T,61) one-hot pitch roll, I (4,16) instrument matrix of the
song, V (N,T) velocity roll, C style class (int), H (N,L) history latents.
"""
from __future__ import annotations

import numpy as np


def prepare_encoder_input_list(X, I, V):
    """[X, I tiled per chunk, V[..., None]]  (vae_definition.py:770-808 with meta_instrument, meta_velocity on)."""
    n = X.shape[0]
    return [X, np.tile(np.expand_dims(I, 0), (n, 1, 1)), np.expand_dims(np.copy(V), 2)]


def prepare_decoder_input(R, output_dim=61, meta_instrument_dim=16, H=None, teacher_force=False, input_length=None, history=True):
    """[Y_start, R, (empty Y), H, I_start, V_start]  (vae_definition.py:816-865)."""
    n = R.shape[0]
    out = [np.zeros((n, output_dim)), R]
    if teacher_force:
        out.append(np.zeros((n, input_length, output_dim)))
    if history:
        if H is None:
            H = np.zeros(R.shape); H[1:] = R[:-1]
        out.append(H)
    out.append(np.zeros((n, meta_instrument_dim)))
    out.append(np.zeros((n,)))
    return out


def prepare_autoencoder_input_and_output_list(X, Y, C, I, V, H, num_classes=2, meta_instrument_dim=16, teacher_force=False,
                                              history=True, silent_weight=1.0, return_sample_weight=False):
    """inputs [X, Y_start, (Y), H, I_start, I, V_start, V]; targets [Y, I, V, C one-hot]; sample weights
    [notes (N,T), style (N,), instrument (N,), velocity (N,)] in the reference's append order (vae_definition.py:880-1045)."""
    n, T = X.shape[0], X.shape[1]
    Vx = np.expand_dims(np.copy(V), 2)
    Cc = np.tile(np.eye(num_classes)[int(C)], (n, 1))
    It = np.tile(np.expand_dims(I, 0), (n, 1, 1))
    inputs = [X, np.zeros((n, Y.shape[2]))]
    if teacher_force:
        inputs.append(Y)
    if history:
        inputs.append(H)
    inputs += [np.zeros((n, meta_instrument_dim)), It, np.zeros((n,)), Vx]
    targets = [Y, It, Vx, Cc]
    if not return_sample_weight:
        return inputs, targets
    w = np.ones((n, T))
    w[np.where(Y[:, :, -1] == 1)] = silent_weight           # :930-933
    return inputs, targets, [w, np.ones((n,)), np.ones((n,)), np.ones((n,))]


def shift_history(z, song_start=None):
    """H[0] = 0, H[i] = z[i-1]  (vae_training.py:791-798)."""
    H = np.zeros_like(z)
    H[1:] = z[:-1]
    if song_start is not None:
        H[np.asarray(song_start, bool)] = 0
    return H
