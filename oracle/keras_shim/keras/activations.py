import torch

from . import backend as K


def get(name):
    if callable(name):
        return name
    if name is None or name == "linear":
        return lambda v: v
    return {"tanh": torch.tanh, "sigmoid": torch.sigmoid, "hard_sigmoid": K.hard_sigmoid, "softmax": K.softmax,
            "relu": torch.relu}[name]
