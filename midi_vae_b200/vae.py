"""VAE: the host-side mirror of the reference's model interface for the hot path.

``VAE().create(**kwargs)`` takes the same keyword arguments as the reference
(vae_definition.py:40-102; call sites vae_training.py:47-109, vae_evaluation.py:488-550) and
exposes ``.encoder``, ``.decoder``, ``.autoencoder`` (and ``.composer_decoder``) objects with
the Keras-Model methods the reference scripts call on them: ``fit``, ``evaluate``,
``predict``, ``train_on_batch``, ``metrics_names``, ``save_weights``, ``load_weights``,
``get_weights``, ``set_weights``, ``reset_states``, ``summary``.  Input / output lists have the
reference's positional layout (vae_definition.py:256-300, 816-865, 880-1045).

All arithmetic happens in libmidivae.so on the GPU; there is no CPU path.  Branches of the
reference that are disabled in its own defaults (settings.py) and are outside the hot path
raise NotImplementedError, loudly.

Extension keywords (not in the reference): ``precision`` ('fp32' | 'bf16'), ``gate_act``,
``dec_cell_variant``, ``decoder_feedback`` ('as_wired' = what the reference graph computes,
'teacher_forced' = what north_star names, 'free_running' for inference), ``max_batch``,
``device``, ``seed``, ``rnn_mode``.
"""
from __future__ import annotations

import warnings
from typing import Dict, List, Optional, Sequence

import numpy as np

from .engine import Engine, EngineConfig, METRIC_KEYS, reference_param_specs

_KERAS_METRIC_NAMES = ["loss", "decoder_loss", "decoder_loss", "decoder_loss", "composer_decoder_loss",
                       "decoder_acc", "decoder_acc", "decoder_acc", "composer_decoder_acc"]
_HISTORY_KEYS = METRIC_KEYS[:9]


class History:
    """keras.callbacks.History look-alike: ``.history`` maps metric name -> one value per epoch."""

    def __init__(self):
        self.history: Dict[str, List[float]] = {k: [] for k in _HISTORY_KEYS}
        self.epoch: List[int] = []


def _pack_onehot(a, n_classes: int, what: str) -> np.ndarray:
    """(…, n_classes) one-hot rolls -> uint8 class indices; anything else is not a roll."""
    a = np.asarray(a)
    if a.shape[-1] != n_classes:
        raise ValueError(f"{what}: last axis must be {n_classes}, got {a.shape}")
    idx = a.argmax(-1)
    ok = (a.max(-1) == 1) & (a.sum(-1) == 1)
    if not bool(ok.all()):
        raise ValueError(f"{what}: rows must be one-hot (the reference feeds one-hot rolls, import_midi.py:243-286)")
    return idx.astype(np.uint8)


def _require_zero(a, what: str):
    if a is not None and np.any(np.asarray(a) != 0):
        raise NotImplementedError(f"{what} must be the all-zero start vector the reference passes (vae_definition.py:820,850,854)")


class _Base:
    def __init__(self, vae: "VAE"):
        self._vae = vae

    # Keras-Model surface shared by the three sub-models
    def reset_states(self):           # nothing in this graph is stateful (vae_training.py:811-812 is a no-op)
        return None

    def summary(self) -> str:
        lines = [f"{type(self).__name__} on cuda:{self._vae.engine.device} ({self._vae.engine.cfg.precision})"]
        total = 0
        for name, shape in self._specs():
            n = int(np.prod(shape)); total += n
            lines.append(f"  {name:<52s} {str(shape):<16s} {n:>10d}")
        lines.append(f"  total params: {total}")
        return "\n".join(lines)

    def _specs(self):
        specs = reference_param_specs(self._vae.engine.cfg)
        first_dec = next(i for i, (n, _) in enumerate(specs) if n.startswith("dec_init/"))
        return {"encoder": specs[:first_dec], "decoder": specs[first_dec:], "autoencoder": specs}[self._part]

    def get_weights(self) -> List[np.ndarray]:
        w = self._vae.engine.get_weights()
        return [w[n] for n, _ in self._specs()]

    def set_weights(self, weights: Sequence[np.ndarray]) -> None:
        specs = self._specs()
        if len(weights) != len(specs):
            raise ValueError(f"expected {len(specs)} weight tensors, got {len(weights)}")
        cur = self._vae.engine.get_weights()
        for (n, _), w in zip(specs, weights):
            cur[n] = np.asarray(w, np.float32)
        self._vae.engine.set_weights(cur)

    def _keras_layout(self):
        from . import keras_names
        return keras_names.layout(self._vae.engine.cfg, self._part)

    def save_weights(self, path: str) -> None:
        """Weights only (optimizer state is not saved, like the reference; vae_training.py:966-978), written as the Keras-2.0.8 HDF5 layout of the
        reference's own checkpoints (one group per layer in model order, ``layer_names`` / ``weight_names`` attributes, float32 datasets) -- the
        reference writes exactly this under its '.pickle' suffix, so files are interchangeable in both directions.  Data only: nothing is pickled."""
        from . import hdf5
        w = self._vae.engine.get_weights()
        hdf5.write_weights(path, [(layer, [(kn, w[en]) for kn, en in ws]) for layer, ws in self._keras_layout()])

    def load_weights(self, path: str, by_name: bool = False) -> None:
        """Keras HDF5 weight files (the reference's shipped ``models/*/*.pickle`` and anything ``save_weights`` wrote).  ``by_name=False`` (what the
        reference calls, vae_evaluation.py:553-559) assigns positionally over the layers that have weights, as Keras 2.0.8 does; ``by_name=True``
        matches layer names.  The file is parsed as data (midi_vae_b200.hdf5); nothing is ever unpickled."""
        from . import hdf5
        with open(path, "rb") as f:
            head = f.read(8)
        if head != hdf5.SIGNATURE:
            raise ValueError(f"{path}: not an HDF5 weight file (the reference writes Keras HDF5 under the '.pickle' suffix; pickled containers are "
                             "not read: unpickling executes code)")
        t = hdf5.read_weights(path)
        mine = [(layer, ws) for layer, ws in self._keras_layout() if ws]
        theirs = [(layer, t["layers"][layer]) for layer in t["layer_names"] if t["layers"][layer]]
        cur = self._vae.engine.get_weights()
        if by_name:
            have = dict(theirs)
            for layer, ws in mine:
                if layer not in have:
                    continue
                if len(have[layer]) != len(ws):
                    raise ValueError(f"{path}: layer {layer} holds {len(have[layer])} weights, this model {len(ws)}")
                for (_, en), (_, arr) in zip(ws, have[layer]):
                    cur[en] = self._fit(en, arr, cur[en], path)
        else:
            if len(theirs) != len(mine):
                raise ValueError(f"{path}: {len(theirs)} layers with weights, this {self._part} has {len(mine)} "
                                 f"(cell_type={self._vae.engine.cfg.cell_type}; the shipped checkpoints are GRU models, settings.py:155)")
            for (layer, ws), (tl, tensors) in zip(mine, theirs):
                if len(ws) != len(tensors):
                    raise ValueError(f"{path}: layer {tl} holds {len(tensors)} weights, layer {layer} of this model {len(ws)}")
                for (_, en), (_, arr) in zip(ws, tensors):
                    cur[en] = self._fit(en, arr, cur[en], path)
        self._vae.engine.set_weights(cur)

    @staticmethod
    def _fit(name, arr, like, path):
        arr = np.asarray(arr, np.float32)
        if arr.shape != like.shape:
            raise ValueError(f"{path}: tensor for {name} has shape {arr.shape}, this model needs {like.shape}")
        return arr


class EncoderModel(_Base):
    """Model([notes_input, meta_instrument_input, meta_velocity_input] -> z)  (vae_definition.py:241-242)."""
    _part = "encoder"

    def predict(self, x, batch_size: int = 32, verbose=False, eps=None):
        X, I, V = x
        vae = self._vae
        P = _pack_onehot(X, vae.engine.cfg.input_dim, "notes_input")
        Ii = _pack_onehot(I, vae.engine.cfg.meta_instrument_dim, "meta_instrument_input")
        Vv = np.asarray(V, np.float32).reshape(P.shape)
        outs = []
        for a in range(0, len(P), min(batch_size, vae.max_batch)):
            b = min(len(P), a + min(batch_size, vae.max_batch))
            e = vae._eps(b - a) if eps is None else np.asarray(eps, np.float32)[a:b]
            outs.append(vae.engine.encode(P[a:b], Ii[a:b], Vv[a:b], e)[0])
        return np.concatenate(outs)


class DecoderModel(_Base):
    """Model([Y0, z, (ground truth), H, I0, V0] -> [Y, I, V])  (vae_definition.py:257-297, 355)."""
    _part = "decoder"

    def predict(self, x, batch_size: int = 32, verbose=False):
        vae = self._vae
        x = list(x)
        Y0 = x.pop(0); z = np.asarray(x.pop(0), np.float32)
        gt = x.pop(0) if vae.teacher_force else None
        H = np.asarray(x.pop(0), np.float32) if vae.engine.cfg.history else None
        I0 = x.pop(0); V0 = x.pop(0)
        _require_zero(Y0, "input_decoder_start"); _require_zero(I0, "input_decoder_meta_instrument_start")
        _require_zero(V0, "input_decoder_meta_velocity_start")
        fb = vae.predict_feedback
        Ys, Is, Vs = [], [], []
        step = min(batch_size, vae.max_batch)
        for a in range(0, len(z), step):
            b = min(len(z), a + step)
            Y, I, V = vae.engine.decode(z[a:b], None if H is None else H[a:b], fb)
            Ys.append(Y); Is.append(I); Vs.append(V[..., None])
        return [np.concatenate(Ys), np.concatenate(Is), np.concatenate(Vs)]


class ComposerDecoderModel(_Base):
    """Model(z -> softmax(z[:, :num_composers]))  (vae_definition.py:730-734): parameter-free, so host-side."""
    _part = "decoder"

    def predict(self, z, batch_size: int = 32, verbose=False):
        z = np.asarray(z, np.float64)[:, :self._vae.engine.cfg.num_composers]
        e = np.exp(z - z.max(-1, keepdims=True))
        return (e / e.sum(-1, keepdims=True)).astype(np.float32)

    def _specs(self):
        return []


class AutoencoderModel(_Base):
    """Model([X, Y0, (Y), H, I0, I, V0, V] -> [Y, I, V, C]) compiled with the reference's loss list
    (vae_definition.py:258-300, 332-441)."""
    _part = "autoencoder"

    @property
    def metrics_names(self) -> List[str]:
        return list(_KERAS_METRIC_NAMES)       # duplicates and all, as Keras 2.0.8 reports them (vae_training.py:172-187)

    # ---- list marshalling -------------------------------------------------------------------
    def _unpack_inputs(self, inputs):
        vae = self._vae
        x = list(inputs)
        X = x.pop(0); Y0 = x.pop(0)
        gt = x.pop(0) if vae.teacher_force else None
        H = np.asarray(x.pop(0), np.float32) if vae.engine.cfg.history else None
        I0 = x.pop(0); I = x.pop(0); V0 = x.pop(0); V = x.pop(0)
        _require_zero(Y0, "input_decoder_start"); _require_zero(I0, "input_decoder_meta_instrument_start")
        _require_zero(V0, "input_decoder_meta_velocity_start")
        cfg = vae.engine.cfg
        P = _pack_onehot(X, cfg.input_dim, "notes_input")
        Ii = _pack_onehot(I, cfg.meta_instrument_dim, "meta_instrument_input")
        Vv = np.asarray(V, np.float32).reshape(P.shape)
        return P, Ii, Vv, H

    def _unpack_targets(self, targets, P, Ii, Vv):
        cfg = self._vae.engine.cfg
        Y, I, V, Cc = targets
        Yp = _pack_onehot(Y, cfg.input_dim, "target notes")
        if not np.array_equal(_pack_onehot(I, cfg.meta_instrument_dim, "target instruments"), Ii):
            raise NotImplementedError("instrument targets must equal the instrument input (vae_definition.py:985-986)")
        if not np.array_equal(np.asarray(V, np.float32).reshape(Vv.shape), Vv):
            raise NotImplementedError("velocity targets must equal the velocity input (vae_definition.py:997-998)")
        style = _pack_onehot(Cc, cfg.num_composers, "style classes")
        return (None if np.array_equal(Yp, P) else Yp), style

    @staticmethod
    def _sample_weight(sample_weight, n, T):
        if sample_weight is None:
            return None
        sw = sample_weight if isinstance(sample_weight, (list, tuple)) else [sample_weight]
        w = np.asarray(sw[0], np.float32)
        if w.shape != (n, T):
            raise ValueError(f"temporal sample weight must be ({n},{T})")
        for extra in sw[1:]:
            if np.any(np.asarray(extra) != 1):
                raise NotImplementedError("per-sample weights other than ones are not used by the reference (vae_definition.py:936-1004)")
        return w

    # ---- Keras surface -------------------------------------------------------------------------
    def train_on_batch(self, x, y, sample_weight=None, eps=None) -> List[float]:
        vae = self._vae
        P, Ii, Vv, H = self._unpack_inputs(x)
        tgt, style = self._unpack_targets(y, P, Ii, Vv)
        w = self._sample_weight(sample_weight, *P.shape)
        e = vae._eps(len(P)) if eps is None else np.asarray(eps, np.float32)
        m = vae.engine.train_on_batch(P, Ii, Vv, style, H, e, w, tgt)
        return [m[k] for k in _HISTORY_KEYS]

    def test_on_batch(self, x, y, sample_weight=None, eps=None) -> List[float]:
        vae = self._vae
        P, Ii, Vv, H = self._unpack_inputs(x)
        tgt, style = self._unpack_targets(y, P, Ii, Vv)
        w = self._sample_weight(sample_weight, *P.shape)
        e = vae._eps(len(P)) if eps is None else np.asarray(eps, np.float32)
        m = vae.engine.evaluate_batch(P, Ii, Vv, style, H, e, w, tgt)
        return [m[k] for k in _HISTORY_KEYS]

    def _loop(self, fn, x, y, batch_size, sample_weight, eps, training=False):
        vae = self._vae
        if batch_size > vae.max_batch:
            # smaller mini-batches than asked for would move the Adam steps (fit) / the temporal-weight normaliser (evaluate): never silently
            msg = (f"batch_size={batch_size} exceeds max_batch={vae.max_batch} of this model: create it with VAE().create(..., max_batch>={batch_size})")
            if training:
                raise ValueError(msg)
            warnings.warn(msg + "; evaluating in mini-batches of max_batch instead")
        P, Ii, Vv, H = self._unpack_inputs(x)
        tgt, style = self._unpack_targets(y, P, Ii, Vv)
        w = self._sample_weight(sample_weight, *P.shape)
        n = len(P)
        step = min(batch_size, vae.max_batch)
        tot = np.zeros(len(_HISTORY_KEYS)); seen = 0
        for a in range(0, n, step):                     # consecutive slices, shuffle=False (vae_training.py:804-809)
            b = min(n, a + step)
            e = vae._eps(b - a) if eps is None else np.asarray(eps, np.float32)[a:b]
            m = fn(P[a:b], Ii[a:b], Vv[a:b], style[a:b], None if H is None else H[a:b], e,
                   None if w is None else w[a:b], None if tgt is None else tgt[a:b])
            tot += np.array([m[k] for k in _HISTORY_KEYS]) * (b - a); seen += b - a
        return tot / max(seen, 1)                       # Keras: batch-size-weighted means

    def fit(self, x, y, epochs: int = 1, batch_size: int = 32, shuffle: bool = False, sample_weight=None, verbose=False, eps=None) -> History:
        if shuffle:
            raise NotImplementedError("the reference always calls fit(shuffle=False) (vae_training.py:807)")
        h = History()
        for ep in range(epochs):
            vals = self._loop(self._vae.engine.train_on_batch, x, y, batch_size, sample_weight, eps, training=True)
            for k, v in zip(_HISTORY_KEYS, vals):
                h.history[k].append(float(v))
            h.epoch.append(ep)
        return h

    def evaluate(self, x, y, batch_size: int = 32, verbose=False, sample_weight=None, eps=None) -> List[float]:
        return [float(v) for v in self._loop(self._vae.engine.evaluate_batch, x, y, batch_size, sample_weight, eps)]

    def predict(self, x, batch_size: int = 32, verbose=False, eps=None):
        vae = self._vae
        P, Ii, Vv, H = self._unpack_inputs(x)
        step = min(batch_size, vae.max_batch)
        outs = [[], [], [], []]
        for a in range(0, len(P), step):
            b = min(len(P), a + step)
            e = vae._eps(b - a) if eps is None else np.asarray(eps, np.float32)[a:b]
            Y, I, V, S, _ = vae.engine.autoencode(P[a:b], Ii[a:b], Vv[a:b], None if H is None else H[a:b], e)
            for o, v in zip(outs, (Y, I, V[..., None], S)):
                o.append(v)
        return [np.concatenate(o) for o in outs]


class VAE(object):
    def create(self, input_dim=64, output_dim=64, use_embedding=False, embedding_dim=0, input_length=16, output_length=16,
               latent_rep_size=256, vae_loss='categorical_crossentropy', optimizer='Adam', activation='sigmoid', lstm_activation='tanh',
               lstm_state_activation='tanh', epsilon_std=1.0, epsilon_factor=0.0, include_composer_decoder=False, num_composers=0,
               composer_weight=1.0, lstm_size=256, cell_type='LSTM', num_layers_encoder=1, num_layers_decoder=1, bidirectional=False,
               decode=True, teacher_force=False, learning_rate=0.001, split_lstm_vector=True, history=True, beta=0.01, prior_mean=0.0,
               prior_std=1.0, decoder_additional_input=False, decoder_additional_input_dim=0, extra_layer=False, meta_instrument=False,
               meta_instrument_dim=0, meta_instrument_length=0, meta_instrument_activation='sigmoid', meta_instrument_weight=1.0,
               signature_decoder=False, signature_dim=0, signature_activation='sigmoid', signature_weight=1.0,
               composer_decoder_at_notes_output=False, composer_decoder_at_notes_weight=1.0,
               composer_decoder_at_notes_activation='softmax', composer_decoder_at_instrument_output=False,
               composer_decoder_at_instrument_weight=1.0, composer_decoder_at_instrument_activation='softmax', meta_velocity=False,
               meta_velocity_length=0, meta_velocity_activation='sigmoid', meta_velocity_weight=1.0, meta_held_notes=False,
               meta_held_notes_length=0, meta_held_notes_activation='softmax', meta_held_notes_weight=1.0, meta_next_notes=False,
               meta_next_notes_output_length=16, meta_next_notes_weight=1.0, meta_next_notes_teacher_force=False,
               activation_before_splitting='tanh',
               # ---- extensions (not in the reference) ----
               precision='fp32', gate_act='hard_sigmoid', dec_cell_variant='recurrentshop_recalled', decoder_feedback='as_wired',
               predict_feedback=None, max_batch=256, device=0, seed=0, rnn_mode='auto'):
        # the reference's own asserts (vae_definition.py:177-208)
        assert num_layers_encoder > 0 and num_layers_decoder > 0
        assert input_length > 0 and output_length > 0 and lstm_size > 0 and latent_rep_size > 0 and beta > 0

        def unsupported(cond, what, where):
            if cond:
                raise NotImplementedError(f"{what} is outside the B200 hot path (disabled in the reference defaults, {where})")
        unsupported(use_embedding, "use_embedding=True", "settings.py:167")
        unsupported(bidirectional, "bidirectional=True", "settings.py:118")
        if cell_type not in ('LSTM', 'GRU'):
            raise NotImplementedError(f"cell_type={cell_type!r}: the B200 path implements the reference's 'GRU' (its shipped default, settings.py:155; "
                                      "step-streamed kernels) and 'LSTM' (north_star's cell; cluster-resident kernels) branches, not SimpleRNN")
        if meta_instrument_length <= 0:
            raise ValueError("meta_instrument_length must be > 0 when meta_instrument=True (settings.py:182 uses 4)")
        unsupported(not meta_instrument or not meta_velocity, "a model without the instrument / velocity streams", "settings.py:180,211")
        unsupported(not include_composer_decoder, "include_composer_decoder=False", "settings.py:133")
        unsupported(meta_held_notes, "meta_held_notes=True", "settings.py:217")
        unsupported(meta_next_notes, "meta_next_notes=True", "settings.py:227")
        unsupported(signature_decoder, "signature_decoder=True", "settings.py:189")
        unsupported(composer_decoder_at_notes_output or composer_decoder_at_instrument_output, "composer decoders at the outputs", "settings.py:195,198")
        unsupported(decoder_additional_input, "decoder_additional_input=True", "settings.py:171")
        unsupported(optimizer != 'Adam', f"optimizer={optimizer!r}", "settings.py:124")
        unsupported(vae_loss != 'categorical_crossentropy', f"vae_loss={vae_loss!r}", "settings.py:125")
        unsupported(activation != 'softmax', f"activation={activation!r}", "settings.py:154")
        unsupported(meta_instrument_activation != 'softmax', f"meta_instrument_activation={meta_instrument_activation!r}", "settings.py:183")
        unsupported(meta_velocity_activation != 'sigmoid', f"meta_velocity_activation={meta_velocity_activation!r}", "settings.py:213")
        unsupported(lstm_activation != 'tanh' or lstm_state_activation != 'tanh' or activation_before_splitting != 'tanh',
                    "non-tanh activations", "settings.py:165-166")
        unsupported(not split_lstm_vector, "split_lstm_vector=False", "settings.py:139")
        unsupported(epsilon_factor > 0, "epsilon_factor > 0", "settings.py:33")
        unsupported(input_dim != output_dim or input_length != output_length, "different input / output shapes", "settings.py:140-153")
        unsupported(meta_velocity_length != input_length, "meta_velocity_length != input_length", "settings.py:212")
        unsupported(not decode, "decode=False", "settings.py:122")

        self.teacher_force = bool(teacher_force)          # changes the input LIST only, as in the reference
        self.epsilon_std = float(epsilon_std)
        self.max_batch = int(max_batch)
        self.predict_feedback = predict_feedback or ("free_running" if decoder_feedback == "teacher_forced" else decoder_feedback)
        self._rng = np.random.default_rng(seed)
        cfg = EngineConfig(
            input_length=input_length, lstm_size=lstm_size, latent_rep_size=latent_rep_size, input_dim=input_dim,
            meta_instrument_dim=meta_instrument_dim, meta_instrument_length=meta_instrument_length, num_composers=num_composers,
            num_layers_encoder=num_layers_encoder, num_layers_decoder=num_layers_decoder, history=bool(history),
            extra_layer=bool(extra_layer), split_lstm_vector=True, gate_act=gate_act, dec_cell_variant=dec_cell_variant,
            decoder_feedback=decoder_feedback, precision=precision, rnn_mode=rnn_mode, max_batch=max_batch, beta=beta, cell_type=cell_type,
            prior_mean=prior_mean, prior_std=prior_std, notes_weight=1.0, meta_instrument_weight=meta_instrument_weight,
            meta_velocity_weight=meta_velocity_weight, composer_weight=composer_weight, learning_rate=learning_rate)
        self.engine = Engine(cfg, device)
        self.engine.set_weights(initial_weights(cfg, seed))
        self.encoder = EncoderModel(self)
        self.decoder = DecoderModel(self)
        self.composer_decoder = ComposerDecoderModel(self)
        self.autoencoder = AutoencoderModel(self)
        self.signature_decoder = None
        return self

    def _eps(self, n: int) -> Optional[np.ndarray]:
        """K.random_normal(stddev=epsilon_std) of the sampling Lambda (vae_definition.py:498-502), drawn on the host."""
        src = getattr(self, "eps_source", None)
        if src is not None:                 # tests replay a recorded sequence of draws (tests/golden/reference_training_loop.npz)
            return np.asarray(src(n), np.float32).reshape(n, self.engine.cfg.latent_rep_size)
        if self.epsilon_std == 0:
            return None
        return (self._rng.standard_normal((n, self.engine.cfg.latent_rep_size)) * self.epsilon_std).astype(np.float32)


def initial_weights(cfg: EngineConfig, seed: int = 42) -> Dict[str, np.ndarray]:
    """Keras default initialisers (Glorot-uniform kernels, orthogonal Keras-LSTM recurrent kernels, zero biases with
    unit forget bias; recurrentshop cells use Dense defaults)."""
    rng = np.random.default_rng(seed)
    H = cfg.lstm_size
    out: Dict[str, np.ndarray] = {}
    for name, shape in reference_param_specs(cfg):
        keras_lstm = name.startswith(("lstm_", "gru_"))        # Keras recurrent layers: orthogonal recurrent kernels
        if name.endswith("/bias"):
            w = np.zeros(shape)
            if name.startswith("lstm_"):                       # unit_forget_bias=True (Keras 2.0.8 LSTM); a GRU has none
                w[H:2 * H] = 1.0
        elif name.endswith("/recurrent_kernel") and keras_lstm:
            a = rng.standard_normal(shape[::-1] if shape[0] < shape[1] else shape)
            q, r = np.linalg.qr(a)
            q = q * np.sign(np.diag(r))
            w = q.T if shape[0] < shape[1] else q
        else:
            limit = np.sqrt(6.0 / (shape[0] + shape[1]))
            w = rng.uniform(-limit, limit, size=shape)
        out[name] = np.ascontiguousarray(w, np.float32).reshape(shape)
    return out
