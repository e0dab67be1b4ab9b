"""Style classifiers on the B200 path: the three evaluator models of the reference --
pitch_classifier.py:89-103, velocity_classifier.py:110-118, instrument_classifier.py:93-103 -- which all are

    Input(shape=(None, input_dim)) -> GRU(lstm_size) x (num_layers - 1, return_sequences) -> GRU(lstm_size) -> Dense(num_classes, softmax)
    compile(Adam(lr), 'categorical_crossentropy', metrics=['accuracy'])

and are combined by vae_evaluation.py:110-117 (``ensemble_prediction``).  ``StyleClassifier`` mirrors the Keras calls those scripts make
(``fit`` / ``evaluate`` / ``predict`` / ``save_weights`` / ``load_weights``) over a libmidivae.so handle in classifier mode
(``mvae_config.model_kind = 1``: ``mvae_cls_train_step_host`` / ``mvae_cls_eval_step_host``).  Weight files are Keras HDF5 with the layer names
of the shipped ``models/*/*_classifier_epoch_*.pickle`` (gru_1, gru_2, dense_1), so those load as they are.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _lib, hdf5
from ._lib import MvaeBatch, MvaeConfig, MvaeMetrics, MvaeParamInfo, check

M_LOSS, M_STYLE_ACC = 0, 8      # mvae_metrics slots (include/midivae.h)


class StyleClassifier:
    """kind 'pitch' (one-hot pitch roll, input_dim classes), 'instrument' (the 4 x 16 one-hot instrument matrix) or 'velocity' (scalar roll)."""

    def __init__(self, kind: str = "pitch", input_dim: int = 61, input_length: int = 64, num_classes: int = 2, lstm_size: int = 256,
                 num_layers: int = 2, learning_rate: float = 1e-5, cell_type: str = "GRU", precision: str = "fp32", max_batch: int = 512,
                 device: int = 0, seed: int = 0):
        if kind not in ("pitch", "instrument", "velocity"):
            raise ValueError("kind must be 'pitch', 'instrument' or 'velocity'")
        self.kind, self.scalar = kind, kind == "velocity"
        self.T, self.D, self.classes, self.H, self.layers = int(input_length), (1 if self.scalar else int(input_dim)), int(num_classes), int(lstm_size), int(num_layers)
        self.max_batch, self.cell = int(max_batch), cell_type
        self.lib = _lib.load()
        c = MvaeConfig()
        check(self.lib.mvae_default_config(C.byref(c)))
        c.input_length, c.lstm_size, c.input_dim = self.T, self.H, (61 if self.scalar else self.D)
        c.num_composers, c.num_layers_encoder, c.latent_rep_size = self.classes, self.layers, max(2, self.classes)
        c.precision, c.max_batch, c.learning_rate = _lib.PRECISION[precision], self.max_batch, learning_rate
        c.cell_type, c.model_kind, c.cls_scalar_input = _lib.CELL_TYPE[cell_type], 1, int(self.scalar)
        self._h = C.c_void_p()
        rc = self.lib.mvae_create(C.byref(c), device, C.byref(self._h))
        if rc != 0:
            msg = self.lib.mvae_last_error(None)
            self._h = None
            raise _lib.MvaeError(f"mvae_create failed ({rc}): {msg.decode() if msg else '?'}")
        n = C.c_int()
        check(self.lib.mvae_param_tensor_count(self._h, C.byref(n)), self._h)
        self._table = []
        for i in range(n.value):
            info = MvaeParamInfo()
            check(self.lib.mvae_param_info_at(self._h, i, C.byref(info)), self._h)
            self._table.append((info.name.decode(), info.rows, info.cols))
        self.set_weights(self.initial_weights(seed))

    # ------------------------------------------------------------------ lifecycle / weights
    def close(self):
        if getattr(self, "_h", None):
            self.lib.mvae_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def weight_specs(self):
        """(name, shape) in the order of the shipped classifier files: gru_k/{kernel, recurrent_kernel, bias}, dense_1/{kernel, bias}."""
        return [(nm, (c,) if nm.endswith("/bias") else (r, c)) for nm, r, c in self._table]

    def initial_weights(self, seed: int = 0) -> Dict[str, np.ndarray]:
        """Keras defaults: glorot_uniform kernels, orthogonal recurrent kernels (per gate block), zero biases (unit forget bias for LSTM)."""
        rng = np.random.default_rng(seed)
        out = {}
        for nm, shape in self.weight_specs():
            if nm.endswith("/bias"):
                w = np.zeros(shape, np.float32)
                if self.cell == "LSTM" and nm.startswith("lstm_"):
                    w[self.H:2 * self.H] = 1.0
            elif nm.endswith("/recurrent_kernel"):
                blocks = []
                for _ in range(shape[1] // self.H):
                    q, r = np.linalg.qr(rng.standard_normal((self.H, self.H)))
                    blocks.append(q * np.sign(np.diag(r)))
                w = np.concatenate(blocks, 1).astype(np.float32)
            else:
                lim = np.sqrt(6.0 / (shape[0] + shape[1]))
                w = rng.uniform(-lim, lim, shape).astype(np.float32)
            out[nm] = w
        return out

    def get_weights(self, grad: bool = False) -> Dict[str, np.ndarray]:
        out = {}
        fn = self.lib.mvae_get_grad if grad else self.lib.mvae_get_param
        for i, (nm, r, c) in enumerate(self._table):
            a = np.empty((r, c), np.float32)
            check(fn(self._h, i, a.ctypes.data_as(C.c_void_p)), self._h)
            out[nm] = a.reshape(-1) if nm.endswith("/bias") else a
        return out

    def get_grads(self):
        return self.get_weights(grad=True)

    def set_weights(self, w: Dict[str, np.ndarray]) -> None:
        for i, (nm, r, c) in enumerate(self._table):
            a = np.ascontiguousarray(np.asarray(w[nm], np.float32).reshape(r, c))
            check(self.lib.mvae_set_param(self._h, i, a.ctypes.data_as(C.c_void_p)), self._h)
        check(self.lib.mvae_commit_params(self._h), self._h)

    def save_weights(self, path: str) -> None:
        w = self.get_weights()
        layers: List = []
        for nm, _ in self.weight_specs():
            layer = nm.split("/")[0]
            if not layers or layers[-1][0] != layer:
                layers.append((layer, []))
            layers[-1][1].append((nm, w[nm]))
        hdf5.write_weights(path, [("input_1", [])] + layers)

    def load_weights(self, path: str) -> None:
        """Keras HDF5 (the shipped models/*/*_classifier_epoch_*.pickle, or save_weights): positional per layer, ':0' suffixes tolerated."""
        tree = hdf5.read_weights(path)
        src = [(n.split(":")[0], a) for layer in tree["layer_names"] for n, a in tree["layers"].get(layer, [])]
        specs = self.weight_specs()
        if len(src) != len(specs):
            raise ValueError(f"{path}: {len(src)} weight tensors, this classifier has {len(specs)}")
        w = {}
        for (nm, shape), (sn, a) in zip(specs, src):
            if tuple(a.shape) != tuple(shape):
                raise ValueError(f"{path}: {sn} has shape {a.shape}, expected {shape} for {nm}")
            w[nm] = a
        self.set_weights(w)

    # ------------------------------------------------------------------ data marshalling
    def _inputs(self, X) -> np.ndarray:
        X = np.asarray(X)
        if self.scalar:
            if X.ndim == 3:
                X = X[..., 0]
            return np.ascontiguousarray(X.reshape(-1, self.T), np.float32)
        if X.ndim == 3:                                  # one-hot rows, as the reference feeds Keras
            if X.shape[-1] != self.D:
                raise ValueError(f"expected one-hot rows of width {self.D}, got {X.shape}")
            if not np.all((X == 0) | (X == 1)) or not np.all(X.sum(-1) <= 1):
                raise ValueError("classifier inputs must be one-hot rolls (at most one 1 per step)")
            X = np.where(X.sum(-1) == 0, 255, X.argmax(-1))     # an all-zero row (unused voice of the instrument matrix) = zero input
        return np.ascontiguousarray(X.reshape(-1, self.T), np.uint8)

    def _labels(self, Y) -> Optional[np.ndarray]:
        if Y is None:
            return None
        Y = np.asarray(Y)
        if Y.ndim == 2:
            Y = Y.argmax(-1)
        return np.ascontiguousarray(Y, np.uint8)

    def _call(self, X, Y, train: bool, probs: Optional[np.ndarray]):
        b = MvaeBatch()
        b.n = len(X)
        if self.scalar:
            b.velocity = X.ctypes.data
        else:
            b.pitch = X.ctypes.data
        if Y is not None:
            b.style = Y.ctypes.data
        m = MvaeMetrics()
        if train:
            check(self.lib.mvae_cls_train_step_host(self._h, C.byref(b), C.byref(m)), self._h)
        else:
            check(self.lib.mvae_cls_eval_step_host(self._h, C.byref(b), C.byref(m), None if probs is None else probs.ctypes.data_as(C.c_void_p)), self._h)
        return float(m.v[M_LOSS]), float(m.v[M_STYLE_ACC])

    # ------------------------------------------------------------------ the Keras surface the classifier scripts use
    def train_on_batch(self, X, Y):
        X, Y = self._inputs(X), self._labels(Y)
        return list(self._call(X, Y, True, None))

    def _loop(self, X, Y, batch_size: int, train: bool):
        X, Y = self._inputs(X), self._labels(Y)
        step = min(int(batch_size), self.max_batch)
        if int(batch_size) > self.max_batch:
            raise ValueError(f"batch_size={batch_size} exceeds max_batch={self.max_batch} of this classifier")
        tot, n = np.zeros(2), len(X)
        for a in range(0, n, step):
            e = min(n, a + step)
            tot += np.array(self._call(X[a:e], Y[a:e], train, None)) * (e - a)
        return tot / max(n, 1)

    def fit(self, X, Y, epochs: int = 1, batch_size: int = 512, shuffle: bool = False, verbose=0):
        """model.fit(X, Y, epochs=1, batch_size, shuffle=False) (pitch_classifier.py:255-262): consecutive mini-batches; history of Keras' names."""
        if shuffle:
            raise NotImplementedError("the classifier scripts call fit(shuffle=False)")
        hist = {"loss": [], "acc": []}
        for _ in range(epochs):
            loss, acc = self._loop(X, Y, batch_size, True)
            hist["loss"].append(float(loss)); hist["acc"].append(float(acc))
        return type("History", (), {"history": hist})()

    def evaluate(self, X, Y, batch_size: int = 512, verbose=0):
        loss, acc = self._loop(X, Y, batch_size, False)
        return [float(loss), float(acc)]

    def predict(self, X, batch_size: int = 512, verbose=0) -> np.ndarray:
        X = self._inputs(X)
        step = min(int(batch_size), self.max_batch)
        out = np.empty((len(X), self.classes), np.float32)
        for a in range(0, len(X), step):
            e = min(len(X), a + step)
            p = np.empty((e - a, self.classes), np.float32)
            self._call(X[a:e], None, False, p)
            out[a:e] = p
        return out

    def reset_states(self):
        pass

    metrics_names = ["loss", "acc"]


def ensemble_prediction(pitch_model: StyleClassifier, instrument_model: StyleClassifier, velocity_model: StyleClassifier, Y, I, V,
                        weights: Sequence[float] = (1.0, 1.0, 1.0)) -> np.ndarray:
    """vae_evaluation.py:110-117: weighted mean of the three classifiers' class probabilities."""
    wp, wi, wv = weights
    return (pitch_model.predict(Y) * wp + instrument_model.predict(I) * wi + velocity_model.predict(V) * wv) / (wp + wi + wv)
