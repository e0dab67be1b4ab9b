"""Keras 2.0.8 initializers used on the path (glorot_uniform, orthogonal, zeros, ones); seeded through numpy."""
import numpy as np
import torch

from .engine import DTYPE

_rng = np.random.RandomState(1337)


def seed(s):
    global _rng
    _rng = np.random.RandomState(s)


def make(kind, shape):
    shape = tuple(int(s) for s in shape)
    if callable(kind):
        return torch.as_tensor(np.asarray(kind(shape)), dtype=DTYPE).clone()
    if kind == "zeros":
        a = np.zeros(shape)
    elif kind == "ones":
        a = np.ones(shape)
    elif kind == "glorot_uniform":
        fan_in, fan_out = (shape[0], shape[1]) if len(shape) == 2 else (shape[0], shape[0])
        lim = np.sqrt(6.0 / (fan_in + fan_out))
        a = _rng.uniform(-lim, lim, size=shape)
    elif kind == "orthogonal":
        flat = _rng.normal(0.0, 1.0, (shape[0], int(np.prod(shape[1:]))))
        u, _, v = np.linalg.svd(flat, full_matrices=False)
        a = (u if u.shape == flat.shape else v).reshape(shape)
    else:
        raise NotImplementedError(f"initializer {kind!r}")
    return torch.as_tensor(a, dtype=DTYPE).clone()
