"""keras.models.Model restatement (Keras 2.0.8 ``engine/training.py`` semantics): compile / fit / evaluate / predict /
train_on_batch with loss_weights, per-output sample weights ('temporal' or 1-D), layer losses (``add_loss``), the
``accuracy`` metric resolution rule and the metrics_names / fit-history naming rules."""
import numpy as np
import torch

from . import objectives
from .engine import DTYPE, KTensor, Layer, as_list, evaluate, ordered_layers


class History:
    def __init__(self):
        self.history = {}


def _weighted_objective(fn, y_true, y_pred, weights):
    """keras.engine.training._weighted_masked_objective (mask = None)."""
    score = fn(y_true, y_pred)
    if weights is not None:
        while score.dim() > weights.dim():
            score = score.mean(dim=-1)
        score = score * weights
        score = score / (weights != 0).to(score.dtype).mean()
    return score.mean()


class Model(Layer):
    def __init__(self, inputs=None, outputs=None, name=None, input=None, output=None):
        super().__init__(name=name)
        inputs = input if inputs is None else inputs
        outputs = output if outputs is None else outputs
        self._single_out = not isinstance(outputs, (list, tuple))
        self.inputs, self.outputs = as_list(inputs), as_list(outputs)
        self.layers = ordered_layers(self.outputs)
        self.built = True
        self.optimizer = None
        # output names = names of the layers that produced the outputs (a layer called once per output repeats its name)
        self.output_names = [t.node.layer.name for t in self.outputs]
        self.input_names = [t.node.layer.name for t in self.inputs]

    # ---- weights
    @property
    def weights(self):
        ws = []
        for l in self.layers:
            ws.extend(l.weights)
        return ws

    @property
    def trainable_weights(self):
        return [w for _, w in self.weights]

    def get_layer(self, name=None, index=None):
        if index is not None:
            return self.layers[index]
        for l in self.layers:
            if l.name == name:
                return l
        raise ValueError(f"No such layer: {name}")

    def weight_layout(self):
        """[(layer_name, weight_name, shape)] in save_weights order (topology.save_weights_to_hdf5_group)."""
        return [(l.name, n, tuple(w.shape)) for l in self.layers for n, w in l.weights]

    def collected_losses(self):
        ls = []
        for l in self.layers:
            ls.extend(l.losses)
            if isinstance(l, Model):
                ls.extend(l.collected_losses())
        return ls

    # ---- use as a layer inside another model (vae_definition.py:355 ``self.decoder(autoencoder_decoder_input_list)``)
    def call(self, inputs, **kwargs):
        ins = as_list(inputs)
        assert len(ins) == len(self.inputs), (self.name, len(ins), len(self.inputs))

        def run(*vals):
            return tuple(evaluate(self.outputs, dict(zip(self.inputs, vals))))
        tup = KTensor(run, ins, None)
        outs = [KTensor((lambda i: (lambda t: t[i]))(i), [tup], o.shape) for i, o in enumerate(self.outputs)]
        return outs[0] if self._single_out else outs

    # ---- numerics
    def _feed(self, x):
        xs = as_list(x)
        assert len(xs) == len(self.inputs), f"{self.name}: expected {len(self.inputs)} input arrays, got {len(xs)}"
        feed = {}
        for t, a in zip(self.inputs, xs):
            a = torch.as_tensor(np.asarray(a), dtype=DTYPE)
            if a.dim() == 1:                   # training._standardize_input_data: 1-D arrays become (N, 1)
                a = a[:, None]
            want = t.shape[1:]
            assert tuple(a.shape[1:]) == tuple(want), f"{self.name}: input {t.name} expected shape (N,)+{want}, got {tuple(a.shape)}"
            feed[t] = a
        return feed, xs[0].shape[0]

    def predict(self, x, batch_size=32, verbose=0):
        feed_all, n = self._feed(x)
        outs = [[] for _ in self.outputs]
        with torch.no_grad():
            for s in range(0, n, batch_size):
                vals = evaluate(self.outputs, {t: a[s:s + batch_size] for t, a in feed_all.items()})
                for o, v in zip(outs, vals):
                    o.append(v)
        res = [torch.cat(o, dim=0).numpy() for o in outs]
        return res[0] if self._single_out else res

    def compile(self, optimizer, loss, loss_weights=None, sample_weight_mode=None, metrics=None):
        self.optimizer = optimizer
        n = len(self.outputs)
        self.loss_functions = [objectives.get(l) for l in (loss if isinstance(loss, (list, tuple)) else [loss] * n)]
        self.loss_weights = list(loss_weights) if loss_weights is not None else [1.0] * n
        modes = sample_weight_mode if isinstance(sample_weight_mode, (list, tuple)) else [sample_weight_mode] * n
        assert len(self.loss_functions) == n and len(self.loss_weights) == n and len(modes) == n
        self.sample_weight_modes = ["temporal" if m == "temporal" else None for m in modes]    # anything else (incl. the string 'None') -> 1-D
        self.metrics = list(metrics or [])
        names = ["loss"]
        if n > 1:
            names += [nm + "_loss" for nm in self.output_names]
        self._metric_fns = []
        for i, (nm, t) in enumerate(zip(self.output_names, self.outputs)):
            for m in self.metrics:
                assert m in ("accuracy", "acc")
                if t.shape[-1] == 1 or self.loss_functions[i] is objectives.binary_crossentropy:
                    fn = objectives.binary_accuracy
                else:
                    fn = objectives.categorical_accuracy
                self._metric_fns.append((i, fn))
                names.append((nm + "_" if n > 1 else "") + "acc")
        self.metrics_names = names

    def _standardize_weights(self, sample_weight, ys):
        n = len(self.outputs)
        sw = as_list(sample_weight) if sample_weight is not None else [None] * n
        assert len(sw) == n
        out = []
        for w, y, mode in zip(sw, ys, self.sample_weight_modes):
            if w is None:
                w = np.ones(y.shape[:2] if mode == "temporal" else y.shape[:1])
            w = torch.as_tensor(np.asarray(w), dtype=DTYPE)
            assert w.dim() == (2 if mode == "temporal" else 1), "sample_weight rank does not match sample_weight_mode"
            assert tuple(w.shape) == tuple(y.shape[:w.dim()])
            out.append(w)
        return out

    def _losses(self, feed, ys, ws):
        tensors = self.outputs + self.collected_losses()
        vals = evaluate(tensors, feed)
        preds, extra = vals[:len(self.outputs)], vals[len(self.outputs):]
        per_output = [_weighted_objective(fn, y, p, w) for fn, y, p, w in zip(self.loss_functions, ys, preds, ws)]
        total = sum(lw * l for lw, l in zip(self.loss_weights, per_output))
        for e in extra:
            total = total + e
        res = [total] + (per_output if len(self.outputs) > 1 else [])
        for i, fn in self._metric_fns:
            res.append(fn(ys[i], preds[i]).mean())
        return res

    def _targets(self, y):
        ys = [torch.as_tensor(np.asarray(a), dtype=DTYPE) for a in as_list(y)]
        ys = [a[:, None] if a.dim() == 1 else a for a in ys]
        assert len(ys) == len(self.outputs)
        for a, t in zip(ys, self.outputs):
            assert tuple(a.shape[1:]) == tuple(t.shape[1:]), f"target shape {tuple(a.shape)} vs output {t.shape}"
        return ys

    def train_on_batch(self, x, y, sample_weight=None):
        feed, _ = self._feed(x)
        ys = self._targets(y)
        ws = self._standardize_weights(sample_weight, ys)
        params = self.trainable_weights
        for p in params:
            p.grad = None
        res = self._losses(feed, ys, ws)
        grads = torch.autograd.grad(res[0], params, allow_unused=True)
        grads = [torch.zeros_like(p) if g is None else g for p, g in zip(params, grads)]
        self.last_grads = {n: g.detach().numpy().copy() for (n, _), g in zip(self.weights, grads)}
        self.optimizer.apply(params, grads)
        out = [float(r) for r in res]
        return out if len(out) > 1 else out[0]

    def test_on_batch(self, x, y, sample_weight=None):
        feed, _ = self._feed(x)
        ys = self._targets(y)
        ws = self._standardize_weights(sample_weight, ys)
        with torch.no_grad():
            return [float(r) for r in self._losses(feed, ys, ws)]

    def _batched(self, fn, x, y, batch_size, sample_weight):
        xs, ys = as_list(x), as_list(y)
        sws = as_list(sample_weight) if sample_weight is not None else None
        n = xs[0].shape[0]
        acc, tot = None, 0
        for s in range(0, n, batch_size):
            sl = slice(s, min(n, s + batch_size))
            r = as_list(fn([a[sl] for a in xs], [a[sl] for a in ys], None if sws is None else [None if w is None else w[sl] for w in sws]))
            k = sl.stop - sl.start
            acc = [v * k for v in r] if acc is None else [a + v * k for a, v in zip(acc, r)]
            tot += k
        return [a / tot for a in acc]

    def evaluate(self, x, y, batch_size=32, verbose=0, sample_weight=None):
        r = self._batched(self.test_on_batch, x, y, batch_size, sample_weight)
        return r if len(r) > 1 else r[0]

    def fit(self, x, y, batch_size=32, epochs=1, verbose=0, shuffle=True, sample_weight=None, **kwargs):
        assert not shuffle, "the reference calls fit(..., shuffle=False) (vae_training.py:804-809); shuffling is not restated"
        labels = self.metrics_names
        dedup = []
        for i, l in enumerate(labels):
            dedup.append(l + "_" + str(labels[:i].count(l) + 1) if labels.count(l) > 1 else l)
        h = History()
        for _ in range(epochs):
            r = self._batched(self.train_on_batch, x, y, batch_size, sample_weight)
            for k, v in zip(dedup, r):
                h.history.setdefault(k, []).append(v)
        return h

    # ---- misc surface the reference scripts touch
    def reset_states(self):
        pass

    def summary(self):
        return "\n".join(f"{l.name:48s} {sum(int(np.prod(w.shape)) for _, w in l.weights)}" for l in self.layers)

    def count_params(self):
        return sum(int(np.prod(w.shape)) for _, w in self.weights)

    def get_weights(self):
        return [w.detach().numpy().copy() for _, w in self.weights]

    def set_weights(self, arrays):
        ws = self.weights
        assert len(ws) == len(arrays), (len(ws), len(arrays))
        for (n, w), a in zip(ws, arrays):
            a = np.asarray(a)
            assert tuple(w.shape) == tuple(a.shape), (n, tuple(w.shape), a.shape)
            with torch.no_grad():
                w.copy_(torch.as_tensor(a, dtype=DTYPE))

    def load_weights(self, path, by_name=False):
        """topology.load_weights_from_hdf5_group: layers with weights are paired POSITIONALLY with the file's layers
        that have weights; tensors inside a layer positionally too.  Reads the file with the repo's pure-Python HDF5 reader."""
        from midi_vae_b200 import hdf5
        t = hdf5.read_weights(path)
        file_layers = [(ln, t["layers"][ln]) for ln in t["layer_names"] if t["layers"][ln]]
        mine = [l for l in self.layers if l.weights]
        assert len(file_layers) == len(mine), f"{path}: {len(file_layers)} layers with weights in the file, {len(mine)} in the model"
        for (ln, tensors), l in zip(file_layers, mine):
            l.set_weights([a for _, a in tensors])
