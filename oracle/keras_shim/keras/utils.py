import numpy as np


def to_categorical(y, num_classes=None):
    """keras.utils.to_categorical (2.0.8): int class vector -> one-hot matrix (n, num_classes)."""
    y = np.array(y, dtype="int").ravel()
    if not num_classes:
        num_classes = np.max(y) + 1
    out = np.zeros((y.shape[0], num_classes))
    out[np.arange(y.shape[0]), y] = 1
    return out
