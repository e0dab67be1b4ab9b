"""recurrentshop.cells restatement (from memory).  Structure -- which Dense layers a cell owns, their shapes and creation
order -- is confirmed by the shipped GRU checkpoints (tests/golden/checkpoint_layout.json: per cell ``dense_k`` kernel (D,3H) +
bias, ``dense_k+1`` kernel (H,2H), ``dense_k+2`` kernel (H,H)).  The gate ORDER inside those matrices is recalled, not confirmed
by shapes; tests/test_reference_pin.py checks it on the shipped trained weights.

LSTM_VARIANT selects between the two conventions SURVEY.md A.3 lists for LSTMCell:
  "recalled": blocks [f|i|c|o], c' = tanh(f*c + i*tanh(.)), h' = o*c'     (what recurrentshop's cells.py is recalled to do)
  "standard": blocks [i|f|c|o], c' = f*c + i*tanh(.),       h' = o*tanh(c')  (Keras LSTM equations)
Both take states in the order (h, c) and return [h', h', c']."""
import torch

from keras import activations
from keras.layers import Dense

from .engine import RNNCell

LSTM_VARIANT = "recalled"
GRU_GATE_ORDER = "zr"          # order of the two gates in the (D,3H) / (H,2H) matrices: "zr" = [z|r|h], "rz" = [r|z|h]
GRU_MIX = "z_keeps_h"          # "z_keeps_h": h' = z*h + (1-z)*hh (Keras 2.0.8);  "z_takes_new": h' = (1-z)*h + z*hh


class LSTMCell(RNNCell):
    def build(self, input_shape):
        H = self.output_dim
        self.kernel = Dense(4 * H, use_bias=self.use_bias)
        self.recurrent_kernel = Dense(4 * H, use_bias=False, kernel_initializer="orthogonal")
        self.kernel.build(input_shape[0]); self.kernel.built = True
        self.recurrent_kernel.build((None, H)); self.recurrent_kernel.built = True
        self.inner = [self.kernel, self.recurrent_kernel]
        self.act = activations.get(self.activation_name)
        self.ract = activations.get(self.recurrent_activation_name)
        self.variant = LSTM_VARIANT

    def step(self, x, h, c):
        H = self.output_dim
        a = x @ self.kernel.kernel + h @ self.recurrent_kernel.kernel
        if self.kernel.bias is not None:
            a = a + self.kernel.bias
        b = [a[:, k * H:(k + 1) * H] for k in range(4)]
        if self.variant == "recalled":
            f, i, g, o = self.ract(b[0]), self.ract(b[1]), self.act(b[2]), self.ract(b[3])
            c2 = self.act(f * c + i * g)
            h2 = o * c2
        else:
            i, f, g, o = self.ract(b[0]), self.ract(b[1]), self.act(b[2]), self.ract(b[3])
            c2 = f * c + i * g
            h2 = o * self.act(c2)
        return h2, h2, c2


class GRUCell(RNNCell):
    def build(self, input_shape):
        H = self.output_dim
        self.kernel = Dense(3 * H, use_bias=self.use_bias)
        self.recurrent_kernel_1 = Dense(2 * H, use_bias=False, kernel_initializer="orthogonal")
        self.recurrent_kernel_2 = Dense(H, use_bias=False, kernel_initializer="orthogonal")
        self.kernel.build(input_shape[0]); self.kernel.built = True
        self.recurrent_kernel_1.build((None, H)); self.recurrent_kernel_1.built = True
        self.recurrent_kernel_2.build((None, H)); self.recurrent_kernel_2.built = True
        self.inner = [self.kernel, self.recurrent_kernel_1, self.recurrent_kernel_2]
        self.act = activations.get(self.activation_name)
        self.ract = activations.get(self.recurrent_activation_name)
        self.gate_order, self.mix = GRU_GATE_ORDER, GRU_MIX

    def step(self, x, h):
        H = self.output_dim
        xa = x @ self.kernel.kernel
        if self.kernel.bias is not None:
            xa = xa + self.kernel.bias
        ra = h @ self.recurrent_kernel_1.kernel
        g0 = self.ract(xa[:, :H] + ra[:, :H])
        g1 = self.ract(xa[:, H:2 * H] + ra[:, H:])
        z, r = (g0, g1) if self.gate_order == "zr" else (g1, g0)
        hh = self.act(xa[:, 2 * H:] + (r * h) @ self.recurrent_kernel_2.kernel)
        h2 = z * h + (1 - z) * hh if self.mix == "z_keeps_h" else (1 - z) * h + z * hh
        return h2, h2


class SimpleRNNCell(RNNCell):
    def __init__(self, *a, **k):
        raise NotImplementedError("SimpleRNNCell: not built at the reference defaults (settings.py:155)")
