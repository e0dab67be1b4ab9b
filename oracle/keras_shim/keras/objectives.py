"""keras.losses / keras.metrics (2.0.8, Theano backend) on evaluated torch tensors."""
import torch

from . import backend as K


def categorical_crossentropy(y_true, y_pred):
    p = y_pred / y_pred.sum(dim=-1, keepdim=True)
    p = torch.clamp(p, K.epsilon(), 1.0 - K.epsilon())
    return -(y_true * torch.log(p)).sum(dim=-1)


def mean_squared_error(y_true, y_pred):
    return torch.mean(torch.square(y_pred - y_true), dim=-1)


mse = MSE = mean_squared_error


def binary_crossentropy(y_true, y_pred):
    p = torch.clamp(y_pred, K.epsilon(), 1.0 - K.epsilon())
    return torch.mean(-(y_true * torch.log(p) + (1 - y_true) * torch.log(1 - p)), dim=-1)


def categorical_accuracy(y_true, y_pred):
    return (y_true.argmax(dim=-1) == y_pred.argmax(dim=-1)).to(y_pred.dtype)


def binary_accuracy(y_true, y_pred):
    return torch.mean((y_true == torch.round(y_pred)).to(y_pred.dtype), dim=-1)


def get(name):
    if callable(name):
        return name
    return {"categorical_crossentropy": categorical_crossentropy, "mse": mse, "mean_squared_error": mse,
            "binary_crossentropy": binary_crossentropy}[name]
