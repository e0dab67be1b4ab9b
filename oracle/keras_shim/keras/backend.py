"""keras.backend restatement (Theano semantics) -- only what vae_definition.py:29-37,498-502 and the layers here call."""
import numpy as np
import torch

from .engine import DTYPE, KTensor, get_uid, reset_uids  # noqa: F401

_EPSILON = 1e-7
_random_normal_hook = None        # tests set this to inject a known epsilon: hook(shape, mean, stddev) -> tensor / ndarray


def epsilon():
    return _EPSILON


def floatx():
    return "float32"


def set_random_normal_hook(fn):
    global _random_normal_hook
    _random_normal_hook = fn


def _lift(fn, x, *a, **k):
    if isinstance(x, KTensor):
        return KTensor(lambda v: fn(v, *a, **k), [x], None)
    r = fn(torch.as_tensor(x, dtype=DTYPE), *a, **k)
    return float(r) if r.dim() == 0 else r


def log(x): return _lift(torch.log, x)
def exp(x): return _lift(torch.exp, x)
def square(x): return _lift(torch.square, x)
def sqrt(x): return _lift(torch.sqrt, x)
def tanh(x): return _lift(torch.tanh, x)
def abs(x): return _lift(torch.abs, x)   # noqa: A001


def _axis_fn(fn):
    def f(x, axis=None, keepdims=False):
        def g(v):
            if axis is None:
                return fn(v)
            return fn(v, dim=axis, keepdim=keepdims)
        return _lift(g, x)
    return f


sum = _axis_fn(torch.sum)     # noqa: A001
mean = _axis_fn(torch.mean)


def shape(x):
    return KTensor(lambda v: tuple(v.shape), [x], None)


def constant(value, dtype=None, shape=None, name=None):
    return torch.as_tensor(value, dtype=DTYPE)


def random_normal(shape, mean=0.0, stddev=1.0, dtype=None, seed=None):
    parents = [s for s in shape if isinstance(s, KTensor)]

    def op(*vals):
        it = iter(vals)
        shp = tuple(int(next(it)) if isinstance(s, KTensor) else int(s) for s in shape)
        if _random_normal_hook is not None:
            return torch.as_tensor(np.asarray(_random_normal_hook(shp, mean, stddev)), dtype=DTYPE)
        return torch.randn(shp, dtype=DTYPE) * stddev + mean
    return KTensor(op, parents, None)


def hard_sigmoid(v):
    # theano.tensor.nnet.hard_sigmoid: clip(0.2 x + 0.5, 0, 1)
    return torch.clamp(0.2 * v + 0.5, 0.0, 1.0)


def softmax(v):
    e = torch.exp(v - v.max(dim=-1, keepdim=True).values)
    return e / e.sum(dim=-1, keepdim=True)


def clear_session():
    reset_uids()
