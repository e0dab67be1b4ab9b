"""recurrentshop.engine restatement: RNNCell (a layer that maps [x, *states] -> [output, *new_states]) and RecurrentModel
in decode mode (from memory)."""
import torch

from keras.engine import KTensor, Layer, as_list, evaluate
from keras.models import Model


class RNNCell(Layer):
    """Base of the recurrentshop cells: ``cell([x, s1, ...]) -> [out, s1', ...]``; weights live in inner Dense layers that
    are created when the cell is first called (which is what gives the shipped checkpoints their dense_k numbering)."""

    def __init__(self, units=None, output_dim=None, activation="tanh", recurrent_activation="hard_sigmoid", use_bias=True, name=None, **kwargs):
        super().__init__(name=name)
        self.output_dim = int(units if units is not None else output_dim)
        self.activation_name, self.recurrent_activation_name, self.use_bias = activation, recurrent_activation, use_bias
        self.inner = []

    @property
    def weights(self):
        return [(f"{self.name}/{n}", w) for l in self.inner for n, w in l.weights]

    def step(self, x, *states):
        raise NotImplementedError

    def call(self, inputs):
        ins = as_list(inputs)
        n_out = len(ins)           # output + one new state per incoming state
        tup = KTensor(lambda *v: tuple(self.step(*v)), ins, None)
        shp = (None, self.output_dim)
        return [KTensor((lambda i: (lambda t: t[i]))(i), [tup], shp) for i in range(n_out)]


class RecurrentModel(Layer):
    """``RecurrentModel(input, output, initial_states, final_states, readout_input, teacher_force, decode, output_length)``.

    decode=True: the 2-D ``input`` is presented unchanged at every one of ``output_length`` steps; the states start from
    ``initial_state`` and are threaded through ``final_states``; the previous step's output (``initial_readout`` before the first
    step; the ground-truth slice when ``teacher_force`` and training) is fed to the ``readout_input`` placeholder; the per-step
    outputs are stacked to (N, output_length, dim).  Whether the readout influences anything depends on the graph the caller
    built -- the reference's graph never connects its readout placeholder (vae_definition.py:545,592,630)."""

    def __init__(self, input, output, initial_states=None, final_states=None, readout_input=None, teacher_force=False, decode=False,
                 output_length=None, return_states=False, state_initializer=None, name=None, **kwargs):
        super().__init__(name=name)
        assert decode and output_length, "only decode mode is restated (the reference passes decode=True, settings.py:123)"
        assert not return_states
        self.input_t, self.output_t = input, output
        self.initial_states, self.final_states = as_list(initial_states or []), as_list(final_states or [])
        assert len(self.initial_states) == len(self.final_states)
        self.readout_input, self.teacher_force, self.output_length = readout_input, teacher_force, int(output_length)
        self.model = Model([input] + self.initial_states, [output] + self.final_states)
        self.built = True
        self.training = True

    @property
    def weights(self):
        return self.model.weights

    def call(self, x, initial_state=None, initial_readout=None, ground_truth=None):
        states = as_list(initial_state or [])
        assert len(states) == len(self.initial_states), "initial_state must provide one tensor per state"
        parents = [x] + states
        has_ro = initial_readout is not None
        has_gt = ground_truth is not None
        if has_ro:
            parents.append(initial_readout)
        if has_gt:
            parents.append(ground_truth)
        ns = len(states)

        def run(*vals):
            xv, sv = vals[0], list(vals[1:1 + ns])
            ro = vals[1 + ns] if has_ro else None
            gt = vals[1 + ns + (1 if has_ro else 0)] if has_gt else None
            outs = []
            for t in range(self.output_length):
                feed = {self.input_t: xv}
                feed.update(dict(zip(self.initial_states, sv)))
                if self.readout_input is not None and ro is not None:
                    feed[self.readout_input] = ro
                res = evaluate([self.output_t] + self.final_states, feed)
                out, sv = res[0], list(res[1:])
                outs.append(out)
                ro = gt[:, t] if (self.teacher_force and self.training and gt is not None) else out
            return torch.stack(outs, dim=1)

        return KTensor(run, parents, (None, self.output_length) + tuple(self.output_t.shape[1:]))
