"""Symbolic tensors, Layer, Input and the functional Model of the Keras-2.0.8 restatement (test infrastructure).

A ``KTensor`` is a lazy node: ``op(*parent_values) -> value``.  Layers are called on KTensors and return KTensors
(static shapes are tracked so that ``build`` can size weights); a ``Model`` evaluates its outputs for a feed of
its Input tensors with torch (float64, autograd on).  Layer ordering inside a Model follows Keras 2.0.8
``Container.__init__`` (depth from the outputs, then discovery order of a post-order DFS), because that order is
what ``save_weights`` writes and ``load_weights(by_name=False)`` relies on (vae_training.py:120-123).
"""
from __future__ import annotations

import re
from collections import OrderedDict

import numpy as np
import torch

DTYPE = torch.float64
_UIDS = {}


def get_uid(prefix):
    _UIDS[prefix] = _UIDS.get(prefix, 0) + 1
    return _UIDS[prefix]


def reset_uids():
    _UIDS.clear()


def _snake(name):
    s = re.sub("(.)([A-Z][a-z0-9]+)", r"\1_\2", name)
    return re.sub("([a-z])([A-Z])", r"\1_\2", s).lower()


class Node:
    def __init__(self, layer, inbound, outputs):
        self.layer, self.inbound, self.outputs = layer, list(inbound), list(outputs)


class KTensor:
    def __init__(self, op, parents, shape=None, name=None):
        self.op, self.parents, self.shape, self.name = op, list(parents), shape, name
        self.node = None          # the layer call that produced this tensor (None for raw backend ops)

    # ---- arithmetic builds new lazy nodes
    def _bin(self, other, fn):
        if isinstance(other, KTensor):
            return KTensor(lambda a, b: fn(a, b), [self, other], self.shape)
        return KTensor(lambda a: fn(a, other), [self], self.shape)

    def __add__(self, o): return self._bin(o, lambda a, b: a + b)
    def __radd__(self, o): return self._bin(o, lambda a, b: b + a)
    def __sub__(self, o): return self._bin(o, lambda a, b: a - b)
    def __rsub__(self, o): return self._bin(o, lambda a, b: b - a)
    def __mul__(self, o): return self._bin(o, lambda a, b: a * b)
    def __rmul__(self, o): return self._bin(o, lambda a, b: b * a)
    def __truediv__(self, o): return self._bin(o, lambda a, b: a / b)
    def __rtruediv__(self, o): return self._bin(o, lambda a, b: b / a)
    def __neg__(self): return KTensor(lambda a: -a, [self], self.shape)

    def __getitem__(self, idx):
        return KTensor(lambda a: a[idx], [self], None)


def evaluate(tensors, feed):
    """Values of ``tensors`` given ``feed`` {KTensor: value}; each node is computed once."""
    memo = {id(k): v for k, v in feed.items()}

    def ev(t):
        stack = [t]
        while stack:
            cur = stack[-1]
            if id(cur) in memo:
                stack.pop()
                continue
            todo = [p for p in cur.parents if id(p) not in memo]
            if todo:
                stack.extend(todo)
                continue
            if cur.op is None:
                raise ValueError(f"placeholder {cur.name!r} was not fed")
            memo[id(cur)] = cur.op(*[memo[id(p)] for p in cur.parents])
            stack.pop()
        return memo[id(t)]

    return [ev(t) for t in tensors]


def as_list(x):
    return list(x) if isinstance(x, (list, tuple)) else [x]


# ------------------------------------------------------------------------------------------------ Layer
class Layer:
    def __init__(self, name=None, trainable=True, **kwargs):
        if not name:
            prefix = _snake(self.__class__.__name__)
            name = prefix + "_" + str(get_uid(prefix))
        self.name = name
        self.trainable = trainable
        self.built = False
        self._weights = []            # [(name, torch tensor requiring grad)]
        self.losses = []              # symbolic scalars registered through add_loss
        self.nodes = []

    # ---- weights
    def add_weight(self, name, shape, initializer="zeros"):
        from . import initializers
        w = initializers.make(initializer, shape)
        w.requires_grad_(True)
        self._weights.append((f"{self.name}/{name}", w))
        return w

    @property
    def weights(self):
        return list(self._weights)

    def get_weights(self):
        return [w.detach().numpy().copy() for _, w in self.weights]

    def set_weights(self, arrays):
        ws = self.weights
        assert len(ws) == len(arrays), (self.name, len(ws), len(arrays))
        for (n, w), a in zip(ws, arrays):
            a = np.asarray(a)
            assert tuple(w.shape) == tuple(a.shape), (n, tuple(w.shape), a.shape)
            with torch.no_grad():
                w.copy_(torch.as_tensor(a, dtype=DTYPE))

    def add_loss(self, loss, inputs=None):
        self.losses.append(loss)

    # ---- calling
    def build(self, input_shape):
        pass

    def compute_output_shape(self, input_shape):
        return input_shape

    def call(self, inputs, **kwargs):
        raise NotImplementedError

    def __call__(self, inputs, **kwargs):
        ins = as_list(inputs)
        shapes = [t.shape for t in ins]
        if not self.built:
            self.build(shapes if isinstance(inputs, (list, tuple)) else shapes[0])
            self.built = True
        out = self.call(inputs, **kwargs)
        outs = as_list(out)
        extra = []
        for v in kwargs.values():          # tensors passed by keyword (recurrentshop: initial_state, initial_readout, ground_truth), in call order
            extra += [x for x in (v if isinstance(v, (list, tuple)) else [v]) if isinstance(x, KTensor)]
        node = Node(self, ins + extra, outs)
        self.nodes.append(node)
        wrapped = []
        for o in outs:
            if o.node is not None:        # a layer that returns its inputs (KLDivergenceLayer): keep identity semantics, new node
                o = KTensor(lambda a: a, [o], o.shape)
            o.node = node
            wrapped.append(o)
        node.outputs = wrapped
        return wrapped if isinstance(out, (list, tuple)) else wrapped[0]


class InputLayer(Layer):
    pass


def Input(shape=None, batch_shape=None, name=None, tensor=None, dtype=None):
    layer = InputLayer(name=name if name else "input_" + str(get_uid("input")))
    if batch_shape is None:
        batch_shape = (None,) + tuple(shape)
    if tensor is not None:
        t = KTensor(lambda: tensor, [], tuple(batch_shape), layer.name)
    else:
        t = KTensor(None, [], tuple(batch_shape), layer.name)
    t.node = Node(layer, [], [t])
    layer.nodes.append(t.node)
    return t


# ------------------------------------------------------------------------------------------------ Container ordering
def ordered_layers(outputs):
    """Keras 2.0.8 ``Container.__init__``: layers by decreasing depth, ties by discovery order of the post-order DFS."""
    layer_indices, nodes_post, seen = OrderedDict(), [], set()

    def build_map(t):
        node = t.node
        if node is None:                      # raw backend op: look through it
            for p in t.parents:
                build_map(p)
            return
        if id(node) in seen:
            return
        seen.add(id(node))
        for p in node.inbound:
            build_map(p)
        if node.layer not in layer_indices:
            layer_indices[node.layer] = len(layer_indices)
        nodes_post.append(node)

    import sys
    sys.setrecursionlimit(max(10000, sys.getrecursionlimit()))
    for t in outputs:
        build_map(t)

    def inbound_nodes(node):
        res, stack = [], list(node.inbound)
        while stack:
            t = stack.pop(0)
            if t.node is None:
                stack = list(t.parents) + stack
            elif t.node is not node:
                res.append(t.node)
        return res

    nodes_depth, layers_depth = {}, {}
    for node in reversed(nodes_post):
        d = nodes_depth.setdefault(id(node), 0)
        d = max(d, layers_depth.get(node.layer, 0))
        layers_depth[node.layer] = d
        nodes_depth[id(node)] = d
        for inn in inbound_nodes(node):
            nodes_depth[id(inn)] = max(d + 1, nodes_depth.get(id(inn), 0))
    by_depth = {}
    for layer, d in layers_depth.items():
        by_depth.setdefault(d, []).append(layer)
    layers = []
    for d in sorted(by_depth, reverse=True):
        layers.extend(sorted(by_depth[d], key=lambda l: layer_indices[l]))
    return layers
