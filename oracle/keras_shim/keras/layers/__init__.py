"""keras.layers restatement (Keras 2.0.8): the layers vae_definition.py:2-8 imports.  Arithmetic follows the
published 2.0.8 sources: Dense; LSTM (kernel blocks [i|f|c|o], recurrent_activation hard_sigmoid, unit_forget_bias);
GRU (blocks [z|r|h], h' = z*h + (1-z)*hh, reset gate applied before the recurrent matmul); Lambda; Concatenate; Add;
Activation.  Bidirectional / Embedding / SimpleRNN / RepeatVector / TimeDistributed are import-only stubs: the reference
imports them but its default configuration (settings.py) never builds them."""
import numpy as np
import torch

from .. import activations
from .. import backend as K  # noqa: F401
from ..engine import DTYPE, Input, InputLayer, KTensor, Layer, as_list  # noqa: F401


class Dense(Layer):
    def __init__(self, units, activation=None, use_bias=True, kernel_initializer="glorot_uniform", bias_initializer="zeros",
                 name=None, **kwargs):
        super().__init__(name=name)
        self.units, self.use_bias = int(units), use_bias
        self.activation = activations.get(activation)
        self.kernel_initializer, self.bias_initializer = kernel_initializer, bias_initializer

    def build(self, input_shape):
        self.kernel = self.add_weight("kernel", (input_shape[-1], self.units), self.kernel_initializer)
        self.bias = self.add_weight("bias", (self.units,), self.bias_initializer) if self.use_bias else None

    def call(self, x):
        def op(v):
            y = v @ self.kernel
            if self.bias is not None:
                y = y + self.bias
            return self.activation(y)
        return KTensor(op, [x], tuple(x.shape[:-1]) + (self.units,))


class Activation(Layer):
    def __init__(self, activation, name=None, **kwargs):
        super().__init__(name=name)
        self.activation = activations.get(activation)

    def call(self, x):
        return KTensor(lambda v: self.activation(v), [x], x.shape)


class Lambda(Layer):
    def __init__(self, function, output_shape=None, name=None, **kwargs):
        super().__init__(name=name)
        self.function, self.output_shape_ = function, output_shape

    def call(self, x):
        out = self.function(x)        # the user function works on symbolic tensors (slicing, K ops)
        assert isinstance(out, KTensor)
        out = KTensor(lambda v: v, [out], None)
        if self.output_shape_ is not None:
            out.shape = (None,) + tuple(self.output_shape_)
        return out


class Concatenate(Layer):
    def __init__(self, axis=-1, name=None, **kwargs):
        super().__init__(name=name)
        self.axis = axis

    def call(self, xs):
        shp = list(xs[0].shape)
        shp[self.axis] = sum(t.shape[self.axis] for t in xs)
        return KTensor(lambda *v: torch.cat(v, dim=self.axis), list(xs), tuple(shp))


class Add(Layer):
    def call(self, xs):
        return KTensor(lambda *v: sum(v[1:], v[0]), list(xs), xs[0].shape)


class Multiply(Layer):
    def call(self, xs):
        def op(*v):
            r = v[0]
            for t in v[1:]:
                r = r * t
            return r
        return KTensor(op, list(xs), xs[0].shape)


def add(xs): return Add()(xs)
def multiply(xs): return Multiply()(xs)
def concatenate(xs, axis=-1): return Concatenate(axis=axis)(xs)


class _Recurrent(Layer):
    GATES = 0

    def __init__(self, units, activation="tanh", recurrent_activation="hard_sigmoid", use_bias=True, return_sequences=False,
                 kernel_initializer="glorot_uniform", recurrent_initializer="orthogonal", bias_initializer="zeros",
                 unit_forget_bias=True, name=None, **kwargs):
        super().__init__(name=name)
        self.units, self.return_sequences, self.use_bias = int(units), return_sequences, use_bias
        self.activation = activations.get(activation)
        self.recurrent_activation = activations.get(recurrent_activation)
        self.kernel_initializer, self.recurrent_initializer, self.bias_initializer = kernel_initializer, recurrent_initializer, bias_initializer
        self.unit_forget_bias = unit_forget_bias

    def build(self, input_shape):
        n = self.GATES * self.units
        self.kernel = self.add_weight("kernel", (input_shape[-1], n), self.kernel_initializer)
        self.recurrent_kernel = self.add_weight("recurrent_kernel", (self.units, n), self.recurrent_initializer)
        self.bias = self.add_weight("bias", (n,), self._bias_init)

    def _bias_init(self, shape):
        return np.zeros(shape)

    def call(self, x):
        T = x.shape[1]
        shp = (None, T, self.units) if self.return_sequences else (None, self.units)
        return KTensor(self._run, [x], shp)


class LSTM(_Recurrent):
    GATES = 4

    def _bias_init(self, shape):
        b = np.zeros(shape)
        if self.unit_forget_bias:
            b[self.units:2 * self.units] = 1.0
        return b

    def _run(self, x):
        H = self.units
        h = torch.zeros(x.shape[0], H, dtype=DTYPE)
        c = torch.zeros(x.shape[0], H, dtype=DTYPE)
        xw = x @ self.kernel + self.bias                 # implementation=0: input projection for all steps at once
        hs = []
        for t in range(x.shape[1]):
            a = xw[:, t] + h @ self.recurrent_kernel
            i = self.recurrent_activation(a[:, :H])
            f = self.recurrent_activation(a[:, H:2 * H])
            c = f * c + i * self.activation(a[:, 2 * H:3 * H])
            o = self.recurrent_activation(a[:, 3 * H:])
            h = o * self.activation(c)
            hs.append(h)
        return torch.stack(hs, dim=1) if self.return_sequences else h


class GRU(_Recurrent):
    GATES = 3

    def _run(self, x):
        H = self.units
        h = torch.zeros(x.shape[0], H, dtype=DTYPE)
        xw = x @ self.kernel + self.bias
        U = self.recurrent_kernel
        hs = []
        for t in range(x.shape[1]):
            z = self.recurrent_activation(xw[:, t, :H] + h @ U[:, :H])
            r = self.recurrent_activation(xw[:, t, H:2 * H] + h @ U[:, H:2 * H])
            hh = self.activation(xw[:, t, 2 * H:] + (r * h) @ U[:, 2 * H:])
            h = z * h + (1 - z) * hh
            hs.append(h)
        return torch.stack(hs, dim=1) if self.return_sequences else h


class _NotOnPath(Layer):
    def __init__(self, *a, **k):
        raise NotImplementedError(f"{self.__class__.__name__}: imported by vae_definition.py but never built at the reference defaults (settings.py)")


class SimpleRNN(_NotOnPath):
    pass


class Bidirectional(_NotOnPath):
    pass


class Embedding(_NotOnPath):
    pass


class RepeatVector(_NotOnPath):
    pass


class TimeDistributed(_NotOnPath):
    pass
