"""CPU oracle for the MIDI-VAE hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu-baseline /
``--impl reference`` legs may import this module.  The product path
(``midi_vae_b200``) never imports it and fails loudly when its CUDA library is
missing.

PARITY PINNED ON THE REFERENCE'S OWN GRAPH CODE, THIRD-PARTY LAYER MATH RESTATED.  The reference
(brunnergino/MIDI-VAE) executes this graph inside Keras 2.0.8 (Theano backend) + recurrentshop
(un-pinned master); neither is vendored under /root/reference nor installed here, and the
reference ships no tests or input/output fixtures.  What pins this file:
  * tests/golden/reference_cfg1.npz -- produced by importing the reference's UNMODIFIED
    vae_definition.py and running VAE.create / prepare_* / predict / evaluate / fit on top of a
    restated slice of Keras 2.0.8 + recurrentshop (oracle/keras_shim, see its README).  This file
    matches those vectors to 1e-9 in float64 (tests/test_reference_pin.py): wiring, positional
    lists, loss composition, metric names, Adam trajectory, both recalled decoder-cell conventions.
  * the shipped HDF5 checkpoints: layer / weight names, shapes and save order reproduced by the
    reference's graph code through the shim (reference_layout.json == checkpoint_layout.json);
    the decoders' untrained first-cell input kernels (=> ``as_wired`` feedback); first decoder
    step of two shipped models on known chords.
What stays un-pinned: the per-layer arithmetic of Keras / recurrentshop is restated from their
published behaviour in BOTH this file and the shim (same author, so not independent), and
recurrentshop's multi-step decode semantics could not be confirmed on the shipped GRU models.
The vectors under tests/golden/cfg1_step.npz are "oracle-derived".

Everything is plain PyTorch on the CPU, written with dense one-hot matmuls as
the reference does (no gather tricks), generic in dtype (float64 for golden
vectors and gradient checks, float32 for the timed CPU baseline).  Gradients
come from autograd; oracle/manual_bptt.py holds the hand-derived backward the
CUDA kernels follow and is itself checked against this file.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field, asdict
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------
# configuration (mirrors the VAE.create kwargs that are live on the hot path,
# vae_definition.py:40-102, defaults from settings.py as listed in SURVEY.md section 5)
# --------------------------------------------------------------------------------------
@dataclass
class OracleConfig:
    input_length: int = 64          # T  (settings.py:108-109,140-142)
    lstm_size: int = 256            # H  (settings.py:110)
    latent_rep_size: int = 256      # L  (settings.py:111)
    input_dim: int = 61             # Dp (settings.py:147-153)
    meta_instrument_dim: int = 16   # Di (settings.py:181)
    meta_instrument_length: int = 4  # Ti (settings.py:182)
    num_composers: int = 2          # C  (settings.py:36-43)
    num_layers_encoder: int = 2     # settings.py:119
    num_layers_decoder: int = 2     # settings.py:120
    beta: float = 0.1               # settings.py:114
    prior_mean: float = 0.0         # settings.py:244
    prior_std: float = 1.0          # settings.py:245
    notes_weight: float = 1.0       # vae_definition.py:336
    meta_instrument_weight: float = 0.1   # settings.py:184
    meta_velocity_weight: float = 1.0     # settings.py:214
    composer_weight: float = 0.1          # settings.py:134
    learning_rate: float = 2e-4     # settings.py:113
    history: bool = True            # settings.py:145
    extra_layer: bool = True        # settings.py:164
    split_lstm_vector: bool = True  # settings.py:139
    # switches that resolve the north_star / reference differences
    gate_act: str = "hard_sigmoid"            # Keras 2.0.8 default; "sigmoid" = north_star wording
    dec_cell_variant: str = "recurrentshop_recalled"   # the reference decoder cell (recurrentshop LSTMCell as recalled); "standard" = Keras LSTM equations
    decoder_feedback: str = "as_wired"        # "as_wired" | "teacher_forced" | "free_running"
    # the reference's shipped default (settings.py:155) is GRU; the CUDA build implements the LSTM branch, the oracle restates both
    # (GRU = groundwork for SURVEY.md 8(f-1); GRUCell conventions as recalled: blocks [z|r|h], h' = (1 - z) h + z hh)
    cell_type: str = "LSTM"                   # "LSTM" | "GRU"

    @property
    def T(self): return self.input_length
    @property
    def H(self): return self.lstm_size
    @property
    def L(self): return self.latent_rep_size


# --------------------------------------------------------------------------------------
# parameter inventory.  Names follow the Keras layer names found in the shipped
# checkpoints (models/*/encoderEpoch*.pickle: gru_1, gru_2, gru_meta_instrument,
# gru_meta_velocity, extra_instrument_after_concat_layer, extra_layer, z_mean,
# z_log_var; decoder: dense_8.. in creation order) with gru -> lstm for the LSTM branch
# (vae_definition.py:457-472).  Decoder tensors get descriptive names; the order is the
# checkpoint's: init-state Denses (notes L1.., instr, vel), then notes cells + output
# Dense, instrument cell + Dense, velocity cell + Dense.
# --------------------------------------------------------------------------------------
def param_specs(cfg: OracleConfig) -> List[Tuple[str, Tuple[int, ...], str]]:
    """(name, shape, init) in canonical (reference checkpoint) order."""
    gru = cfg.cell_type == "GRU"
    H, L, G = cfg.H, cfg.L, (3 if gru else 4) * cfg.H
    Dp, Di = cfg.input_dim, cfg.meta_instrument_dim
    s: List[Tuple[str, Tuple[int, ...], str]] = []
    pre = "gru" if gru else "lstm"

    def keras_lstm(name, D):
        s.append((f"{name}/kernel", (D, G), "glorot"))
        s.append((f"{name}/recurrent_kernel", (H, G), "orthogonal"))
        s.append((f"{name}/bias", (G,), "zeros" if gru else "lstm_bias"))

    for k in range(1, cfg.num_layers_encoder + 1):           # vae_definition.py:455-460
        keras_lstm(f"{pre}_{k}", Dp if k == 1 else H)
    keras_lstm(f"{pre}_meta_instrument", Di)                 # :464-468
    keras_lstm(f"{pre}_meta_velocity", 1)                    # :470-474
    s.append(("extra_instrument_after_concat_layer/kernel", (3 * H, H), "glorot"))  # :483-484
    s.append(("extra_instrument_after_concat_layer/bias", (H,), "zeros"))
    if cfg.extra_layer:
        s.append(("extra_layer/kernel", (H, H), "glorot"))   # :486-487
        s.append(("extra_layer/bias", (H,), "zeros"))
    half = H // 2 if cfg.split_lstm_vector else H            # :489-496
    rest = H - half if cfg.split_lstm_vector else H
    s.append(("z_mean/kernel", (half, L), "glorot"))         # :506
    s.append(("z_mean/bias", (L,), "zeros"))
    s.append(("z_log_var/kernel", (rest, L), "glorot"))      # :507
    s.append(("z_log_var/bias", (L,), "zeros"))

    Q = 2 * L if cfg.history else L                          # :548-551
    def init_dense(name):
        for j in ((1,) if gru else (1, 2)):                  # :563-568 (state 1, state 2; a GRU cell has one state)
            s.append((f"dec_init/{name}_s{j}/kernel", (Q, H), "glorot"))
            s.append((f"dec_init/{name}_s{j}/bias", (H,), "zeros"))
    for k in range(1, cfg.num_layers_decoder + 1):
        init_dense(f"notes_l{k}")
    init_dense("instr")                                      # :599-604
    init_dense("vel")                                        # :637-642

    def rs_cell(name, D):                                    # recurrentshop LSTMCell: Dense(4H,bias) on x, Dense(4H,no bias) on h
        s.append((f"{name}/kernel", (D, G), "glorot"))
        s.append((f"{name}/bias", (G,), "zeros"))
        if gru:                                              # GRUCell: Dense(3H,bias) on x, Dense(2H) on h for z, r, Dense(H) on r*h (shipped shapes)
            s.append((f"{name}/recurrent_kernel_1", (H, 2 * H), "glorot"))
            s.append((f"{name}/recurrent_kernel_2", (H, H), "glorot"))
        else:
            s.append((f"{name}/recurrent_kernel", (H, G), "glorot"))
    for k in range(1, cfg.num_layers_decoder + 1):           # :533-540
        rs_cell(f"notes/cell_{k}", Dp if k == 1 else H)
    s.append(("notes/out/kernel", (H, Dp), "glorot"))        # :542
    s.append(("notes/out/bias", (Dp,), "zeros"))
    rs_cell("meta_instrument/cell", Di)                      # :583-590
    s.append(("meta_instrument/out/kernel", (H, Di), "glorot"))  # :593
    s.append(("meta_instrument/out/bias", (Di,), "zeros"))
    rs_cell("meta_velocity/cell", 1)                         # :621-628
    s.append(("meta_velocity/out/kernel", (H, 1), "glorot"))  # :631
    s.append(("meta_velocity/out/bias", (1,), "zeros"))
    return s


def param_count(cfg: OracleConfig) -> int:
    return sum(int(np.prod(shape)) for _, shape, _ in param_specs(cfg))


def init_params(cfg: OracleConfig, seed: int = 42, dtype=torch.float64) -> Dict[str, Tensor]:
    """Keras default initialisers: Glorot-uniform kernels, orthogonal recurrent kernels
    (Keras LSTM), zero biases with unit forget bias (Keras 2.0.8 LSTM unit_forget_bias=True);
    recurrentshop cells use Dense defaults (Glorot-uniform, zero bias)."""
    rng = np.random.default_rng(seed)
    H = cfg.H
    out: Dict[str, Tensor] = {}
    for name, shape, init in param_specs(cfg):
        if init == "glorot":
            limit = math.sqrt(6.0 / (shape[0] + shape[1]))
            w = rng.uniform(-limit, limit, size=shape)
        elif init == "orthogonal":
            a = rng.standard_normal(shape if shape[0] >= shape[1] else shape[::-1])
            q, r = np.linalg.qr(a)
            q = q * np.sign(np.diag(r))
            w = q if shape[0] >= shape[1] else q.T
            w = np.ascontiguousarray(w[: shape[0], : shape[1]])
        elif init == "lstm_bias":
            w = np.zeros(shape)
            w[H:2 * H] = 1.0
        else:
            w = np.zeros(shape)
        out[name] = torch.tensor(w, dtype=dtype)
    return out


# --------------------------------------------------------------------------------------
# cells
# --------------------------------------------------------------------------------------
def hard_sigmoid(x: Tensor) -> Tensor:
    """Keras/Theano hard_sigmoid: clip(0.2 x + 0.5, 0, 1)."""
    return torch.clamp(0.2 * x + 0.5, 0.0, 1.0)


def _gate(cfg: OracleConfig):
    return hard_sigmoid if cfg.gate_act == "hard_sigmoid" else torch.sigmoid


def lstm_step(cfg: OracleConfig, a: Tensor, c: Tensor, variant: str = "standard") -> Tuple[Tensor, Tensor]:
    """One LSTM cell update from pre-activations a = x W + h U + b   (B,4H).

    standard (Keras 2.0.8 LSTM, blocks [i|f|c|o]):  c' = f*c + i*tanh(g);  h' = o*tanh(c').
    recurrentshop_recalled (blocks [f|i|c|o], external, from memory):  c' = tanh(f*c + i*tanh(g)); h' = o*c'.
    """
    H = cfg.H
    act = _gate(cfg)
    if variant == "standard":
        i = act(a[:, 0:H]); f = act(a[:, H:2 * H]); g = torch.tanh(a[:, 2 * H:3 * H]); o = act(a[:, 3 * H:4 * H])
        c_new = f * c + i * g
        h_new = o * torch.tanh(c_new)
    elif variant == "recurrentshop_recalled":
        f = act(a[:, 0:H]); i = act(a[:, H:2 * H]); g = torch.tanh(a[:, 2 * H:3 * H]); o = act(a[:, 3 * H:4 * H])
        c_new = torch.tanh(f * c + i * g)
        h_new = o * c_new
    else:
        raise ValueError(variant)
    return h_new, c_new


def keras_lstm(cfg: OracleConfig, p: Dict[str, Tensor], name: str, x: Tensor, return_sequences: bool) -> Tensor:
    """keras.layers.LSTM(H, return_sequences=...) with zero initial state (vae_definition.py:457-472)."""
    B, T, _ = x.shape
    W, U, b = p[f"{name}/kernel"], p[f"{name}/recurrent_kernel"], p[f"{name}/bias"]
    h = x.new_zeros(B, cfg.H); c = x.new_zeros(B, cfg.H)
    xw = x @ W + b                       # implementation=0: input projection precomputed
    hs = []
    for t in range(T):
        h, c = lstm_step(cfg, xw[:, t] + h @ U, c, "standard")
        hs.append(h)
    return torch.stack(hs, 1) if return_sequences else h


def keras_gru(cfg: OracleConfig, p: Dict[str, Tensor], name: str, x: Tensor, return_sequences: bool) -> Tensor:
    """keras.layers.GRU (2.0.8): blocks [z|r|h]; z, r = gate(x W + h U + b); hh = tanh(x W_h + (r*h) U_h + b_h); h' = z*h + (1-z)*hh."""
    B, T, _ = x.shape
    H = cfg.H
    W, U, b = p[f"{name}/kernel"], p[f"{name}/recurrent_kernel"], p[f"{name}/bias"]
    act = _gate(cfg)
    h = x.new_zeros(B, H)
    xw = x @ W + b
    hs = []
    for t in range(T):
        z = act(xw[:, t, :H] + h @ U[:, :H])
        r = act(xw[:, t, H:2 * H] + h @ U[:, H:2 * H])
        hh = torch.tanh(xw[:, t, 2 * H:] + (r * h) @ U[:, 2 * H:])
        h = z * h + (1 - z) * hh
        hs.append(h)
    return torch.stack(hs, 1) if return_sequences else h


# --------------------------------------------------------------------------------------
# encoder  (vae_definition.py:443-516)
# --------------------------------------------------------------------------------------
def encoder_heads(cfg: OracleConfig, p: Dict[str, Tensor], X: Tensor, I: Tensor, V: Tensor) -> Tuple[Tensor, Tensor]:
    """X (B,T,61) one-hot, I (B,4,16) one-hot, V (B,T,1)  ->  (z_mean, z_log_var)  each (B,L)."""
    gru = cfg.cell_type == "GRU"
    rnn, pre = (keras_gru, "gru") if gru else (keras_lstm, "lstm")
    h = X
    for k in range(1, cfg.num_layers_encoder):
        h = rnn(cfg, p, f"{pre}_{k}", h, True)
    h = rnn(cfg, p, f"{pre}_{cfg.num_layers_encoder}", h, False)
    m_i = rnn(cfg, p, f"{pre}_meta_instrument", I, False)
    m_v = rnn(cfg, p, f"{pre}_meta_velocity", V, False)
    u = torch.cat([h, m_i, m_v], dim=1)                                   # :468,:474
    a = torch.tanh(u @ p["extra_instrument_after_concat_layer/kernel"] + p["extra_instrument_after_concat_layer/bias"])
    if cfg.extra_layer:
        a = torch.tanh(a @ p["extra_layer/kernel"] + p["extra_layer/bias"])
    if cfg.split_lstm_vector:
        half = cfg.H // 2
        h1, h2 = a[:, :half], a[:, half:]
    else:
        h1 = h2 = a
    mu = h1 @ p["z_mean/kernel"] + p["z_mean/bias"]
    logvar = h2 @ p["z_log_var/kernel"] + p["z_log_var/bias"]
    return mu, logvar


def kl_term(cfg: OracleConfig, mu: Tensor, logvar: Tensor) -> Tensor:
    """KLDivergenceLayer.call (vae_definition.py:29-37): beta * mean_b(-0.5 * sum_j(...))."""
    prior_log_var = math.log(cfg.prior_std) * 2
    prior_var = cfg.prior_std ** 2
    kl_batch = cfg.beta * (-0.5 * torch.sum(1 + logvar - prior_log_var - ((mu - cfg.prior_mean) ** 2 + torch.exp(logvar)) / prior_var, dim=1))
    return kl_batch.mean()


def sample_z(mu: Tensor, logvar: Tensor, eps: Tensor) -> Tensor:
    """sampling Lambda (vae_definition.py:498-502); eps ~ N(0, epsilon_std^2) is an explicit input."""
    return mu + torch.exp(logvar / 2) * eps


def encode(cfg, p, X, I, V, eps=None):
    mu, logvar = encoder_heads(cfg, p, X, I, V)
    z = mu if eps is None else sample_z(mu, logvar, eps)
    return z, mu, logvar


# --------------------------------------------------------------------------------------
# decoder  (vae_definition.py:519-645)
# --------------------------------------------------------------------------------------
def _init_states(cfg, p, q, name):
    s1 = torch.tanh(q @ p[f"dec_init/{name}_s1/kernel"] + p[f"dec_init/{name}_s1/bias"])
    if cfg.cell_type == "GRU":
        return s1, None
    s2 = torch.tanh(q @ p[f"dec_init/{name}_s2/kernel"] + p[f"dec_init/{name}_s2/bias"])
    return s1, s2            # (state1, state2) = (h, c)


def _rs_cell(cfg, p, name, x, h, c):
    if cfg.cell_type == "GRU":
        # recurrentshop GRUCell as recalled (structure = the shipped checkpoints): blocks [z|r|h]; h' = (1 - z) h + z hh  -- note the mix is
        # the opposite of Keras' GRU layer; the first decoder step of the shipped models decodes correctly only with this one
        H = cfg.H
        act = _gate(cfg)
        xa = x @ p[f"{name}/kernel"] + p[f"{name}/bias"]
        ra = h @ p[f"{name}/recurrent_kernel_1"]
        z = act(xa[:, :H] + ra[:, :H])
        r = act(xa[:, H:2 * H] + ra[:, H:])
        hh = torch.tanh(xa[:, 2 * H:] + (r * h) @ p[f"{name}/recurrent_kernel_2"])
        return (1 - z) * h + z * hh, None
    a = x @ p[f"{name}/kernel"] + p[f"{name}/bias"] + h @ p[f"{name}/recurrent_kernel"]
    return lstm_step(cfg, a, c, cfg.dec_cell_variant)


def decode(cfg: OracleConfig, p: Dict[str, Tensor], z: Tensor, hist: Optional[Tensor],
           Y: Optional[Tensor] = None, I: Optional[Tensor] = None, V: Optional[Tensor] = None,
           feedback: Optional[str] = None,
           Y0: Optional[Tensor] = None, I0: Optional[Tensor] = None, V0: Optional[Tensor] = None
           ) -> Tuple[Tensor, Tensor, Tensor]:
    """-> (Yhat (B,T,61) softmax, Ihat (B,4,16) softmax, Vhat (B,T,1) sigmoid).

    feedback (SURVEY.md section 8(a) row D-fb):
      as_wired        x_t = start vector for every t (the reference's own graph: readout unconnected)
      teacher_forced  x_0 = start, x_t = target_{t-1}      (needs Y, I, V)
      free_running    x_0 = start, x_t = prediction_{t-1}
    """
    fb = feedback or cfg.decoder_feedback
    B = z.shape[0]
    T, Ti = cfg.T, cfg.meta_instrument_length
    q = torch.cat([z, hist], dim=1) if cfg.history else z
    zeros = lambda d: z.new_zeros(B, d)
    Y0 = zeros(cfg.input_dim) if Y0 is None else Y0
    I0 = zeros(cfg.meta_instrument_dim) if I0 is None else I0
    V0 = zeros(1) if V0 is None else V0.reshape(B, 1)

    def run(cells, out_name, steps, start, target, out_act):
        states = [list(_init_states(cfg, p, q, nm)) for nm, _ in cells]
        outs = []
        x = start
        for t in range(steps):
            if fb == "as_wired":
                x = start
            elif fb == "teacher_forced":
                x = start if t == 0 else target[:, t - 1]
            elif fb == "free_running":
                x = start if t == 0 else outs[-1]
            else:
                raise ValueError(fb)
            inp = x
            for li, (_, cname) in enumerate(cells):
                h, c = _rs_cell(cfg, p, cname, inp, states[li][0], states[li][1])
                states[li] = [h, c]
                inp = h
            logits = inp @ p[f"{out_name}/kernel"] + p[f"{out_name}/bias"]
            outs.append(out_act(logits))
        return torch.stack(outs, 1)

    notes_cells = [(f"notes_l{k}", f"notes/cell_{k}") for k in range(1, cfg.num_layers_decoder + 1)]
    Yh = run(notes_cells, "notes/out", T, Y0, Y, lambda a: torch.softmax(a, dim=-1))
    Ih = run([("instr", "meta_instrument/cell")], "meta_instrument/out", Ti, I0, I, lambda a: torch.softmax(a, dim=-1))
    Vh = run([("vel", "meta_velocity/cell")], "meta_velocity/out", T, V0, V, torch.sigmoid)
    return Yh, Ih, Vh


def style_head(cfg: OracleConfig, z: Tensor) -> Tensor:
    """_build_composer_decoder (vae_definition.py:730-734): softmax(z[:, :C]); no parameters."""
    return torch.softmax(z[:, :cfg.num_composers], dim=-1)


# --------------------------------------------------------------------------------------
# Keras 2.0.8 losses / metrics (weighted objectives)
# --------------------------------------------------------------------------------------
_EPS = 1e-7


def categorical_crossentropy(y: Tensor, p: Tensor) -> Tensor:
    """Theano-backend K.categorical_crossentropy: renormalise, clip to [eps, 1-eps], -sum y log p."""
    p = p / p.sum(dim=-1, keepdim=True)
    p = torch.clamp(p, _EPS, 1.0 - _EPS)
    return -(y * torch.log(p)).sum(dim=-1)


def _weighted(score: Tensor, w: Optional[Tensor]) -> Tensor:
    """keras.engine.training._weighted_masked_objective: reduce to weight ndim, * w / mean(w != 0), mean."""
    if w is None:
        w = score.new_ones(score.shape[0])
    while score.dim() > w.dim():
        score = score.mean(dim=-1)
    score = score * w / (w != 0).to(score.dtype).mean()
    return score.mean()


def losses_and_metrics(cfg: OracleConfig, outs, targets, kl: Tensor, sample_weight=None) -> Dict[str, Tensor]:
    """outs = (Yh, Ih, Vh, Ch); targets = (Y, I, V, C); sample_weight = (w_notes (B,T), w_instr, w_vel, w_style) or None.
    Returns the dict keyed like hist.history (vae_training.py:817-853); per-output losses are unweighted by
    loss_weights, 'loss' is the weighted total incl. KL."""
    Yh, Ih, Vh, Ch = outs
    Y, I, V, C = targets
    sw = sample_weight or (None, None, None, None)
    w_notes = sw[0] if sw[0] is not None else Yh.new_ones(Yh.shape[:2])
    l_notes = _weighted(categorical_crossentropy(Y, Yh), w_notes)
    l_instr = _weighted(categorical_crossentropy(I, Ih), sw[1])
    l_vel = _weighted(((Vh - V) ** 2).mean(dim=-1), sw[2])
    l_style = _weighted(categorical_crossentropy(C, Ch), sw[3])
    total = (cfg.notes_weight * l_notes + cfg.meta_instrument_weight * l_instr
             + cfg.meta_velocity_weight * l_vel + cfg.composer_weight * l_style + kl)
    dt = Yh.dtype
    # Keras 2.0.8 metrics go through _masked_objective: a plain mean, sample weights do NOT enter (weighted_metrics arrived in 2.0.9);
    # pinned by tests/test_reference_pin.py on the reference-executed variant with silent_weight != 1
    acc_notes = (Yh.argmax(-1) == Y.argmax(-1)).to(dt).mean()
    acc_instr = (Ih.argmax(-1) == I.argmax(-1)).to(dt).mean()
    acc_vel = (torch.round(Vh) == V).to(dt).mean()
    acc_style = (Ch.argmax(-1) == C.argmax(-1)).to(dt).mean()
    return {
        "loss": total, "decoder_loss_1": l_notes, "decoder_loss_2": l_instr, "decoder_loss_3": l_vel,
        "composer_decoder_loss": l_style, "decoder_acc_1": acc_notes, "decoder_acc_2": acc_instr,
        "decoder_acc_3": acc_vel, "composer_decoder_acc": acc_style, "kl": kl,
    }


METRICS_NAMES = ["loss", "decoder_loss_1", "decoder_loss_2", "decoder_loss_3", "composer_decoder_loss",
                 "decoder_acc_1", "decoder_acc_2", "decoder_acc_3", "composer_decoder_acc"]


# --------------------------------------------------------------------------------------
# autoencoder forward / train step
# --------------------------------------------------------------------------------------
def autoencoder_forward(cfg, p, X, I, V, hist, eps, Y=None, feedback=None):
    """autoencoder = encoder -> decoder(+style head) (vae_definition.py:355-357,391-397,436)."""
    Y = X if Y is None else Y
    mu, logvar = encoder_heads(cfg, p, X, I, V)
    kl = kl_term(cfg, mu, logvar)
    z = sample_z(mu, logvar, eps)
    Yh, Ih, Vh = decode(cfg, p, z, hist, Y, I, V, feedback)
    Ch = style_head(cfg, z)
    return (Yh, Ih, Vh, Ch), kl, (z, mu, logvar)


def one_hot(idx: Tensor, n: int, dtype) -> Tensor:
    return torch.nn.functional.one_hot(idx.long(), n).to(dtype)


def batch_from_packed(cfg, pitch_idx, instr_idx, velocity, style_idx, dtype, tgt_idx=None):
    """Expand the packed boundary layout (u8 indices) into the reference's dense tensors."""
    X = one_hot(torch.as_tensor(pitch_idx), cfg.input_dim, dtype)
    Y = X if tgt_idx is None else one_hot(torch.as_tensor(tgt_idx), cfg.input_dim, dtype)
    I = one_hot(torch.as_tensor(instr_idx), cfg.meta_instrument_dim, dtype)
    V = torch.as_tensor(velocity, dtype=dtype).unsqueeze(-1)
    C = one_hot(torch.as_tensor(style_idx), cfg.num_composers, dtype)
    return X, Y, I, V, C


class KerasAdam:
    """keras.optimizers.Adam (2.0.8): lr_t = lr*sqrt(1-b2^t)/(1-b1^t); p -= lr_t*m/(sqrt(v)+eps)."""

    def __init__(self, params: Dict[str, Tensor], lr=2e-4, beta_1=0.9, beta_2=0.999, epsilon=1e-8):
        self.lr, self.b1, self.b2, self.eps = lr, beta_1, beta_2, epsilon
        self.t = 0
        self.m = {k: torch.zeros_like(v) for k, v in params.items()}
        self.v = {k: torch.zeros_like(v) for k, v in params.items()}

    def step(self, params: Dict[str, Tensor], grads: Dict[str, Tensor]) -> None:
        self.t += 1
        lr_t = self.lr * math.sqrt(1.0 - self.b2 ** self.t) / (1.0 - self.b1 ** self.t)
        with torch.no_grad():
            for k, pk in params.items():
                g = grads[k]
                self.m[k].mul_(self.b1).add_(g, alpha=1 - self.b1)
                self.v[k].mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
                pk.sub_(lr_t * self.m[k] / (self.v[k].sqrt() + self.eps))


def loss_and_grads(cfg, p, X, I, V, C, hist, eps, Y=None, sample_weight=None, feedback=None):
    """Forward + autograd backward; returns (metrics dict of python floats, grads dict, aux)."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in p.items()}
    outs, kl, aux = autoencoder_forward(cfg, leaves, X, I, V, hist, eps, Y, feedback)
    m = losses_and_metrics(cfg, outs, (X if Y is None else Y, I, V, C), kl, sample_weight)
    names = list(leaves.keys())
    gl = torch.autograd.grad(m["loss"], [leaves[k] for k in names], allow_unused=True)
    grads = {k: (g if g is not None else torch.zeros_like(leaves[k])) for k, g in zip(names, gl)}
    return {k: float(v.detach()) for k, v in m.items()}, grads, (tuple(o.detach() for o in outs), tuple(a.detach() for a in aux))


def train_on_batch(cfg, p, opt: KerasAdam, X, I, V, C, hist, eps, Y=None, sample_weight=None, feedback=None):
    """One mini-batch of autoencoder.fit (vae_training.py:804-809): fwd + bwd + Keras Adam, in place on p."""
    metrics, grads, aux = loss_and_grads(cfg, p, X, I, V, C, hist, eps, Y, sample_weight, feedback)
    opt.step(p, grads)
    return metrics, aux


def evaluate_batch(cfg, p, X, I, V, C, hist, eps, Y=None, sample_weight=None, feedback=None):
    with torch.no_grad():
        outs, kl, aux = autoencoder_forward(cfg, p, X, I, V, hist, eps, Y, feedback)
        m = losses_and_metrics(cfg, outs, (X if Y is None else Y, I, V, C), kl, sample_weight)
    return {k: float(v) for k, v in m.items()}, outs, aux


# --------------------------------------------------------------------------------------
# inference: encode -> swap -> decode  (vae_evaluation.py:2180-2181, 2448-2550; appendix A.5)
# --------------------------------------------------------------------------------------
def shift_history(z: Tensor, song_start: Optional[np.ndarray] = None) -> Tensor:
    """H[0] = 0, H[i] = z[i-1] within each song (vae_training.py:791-798)."""
    Hs = torch.zeros_like(z)
    Hs[1:] = z[:-1]
    if song_start is not None:
        Hs[torch.as_tensor(np.asarray(song_start), dtype=torch.bool)] = 0
    return Hs


def style_transfer(cfg, p, X, I, V, c_from=0, c_to=1, song_start=None, feedback=None):
    """Batched form of the per-chunk loop at vae_evaluation.py:2471-2550 (eps = 0 => z = mu)."""
    with torch.no_grad():
        mu, _ = encoder_heads(cfg, p, X, I, V)
        z_sw = mu.clone()
        z_sw[:, c_from], z_sw[:, c_to] = mu[:, c_to], mu[:, c_from]      # :2474-2478
        H_sw = shift_history(z_sw, song_start)                            # :2460,2481,2550
        fb = feedback or ("free_running" if cfg.decoder_feedback == "teacher_forced" else cfg.decoder_feedback)
        Yh, Ih, Vh = decode(cfg, p, z_sw, H_sw, None, None, None, fb)
        pitch = Yh.argmax(-1)             # 60 == silent (vae_definition.py:1090-1091)
        instr = Ih.argmax(-1)
    return {"z": mu, "z_sw": z_sw, "H_sw": H_sw, "Yh": Yh, "Ih": Ih, "Vh": Vh, "pitch": pitch, "instr": instr}


def top2_margin(prob: Tensor) -> Tensor:
    """top-1 minus top-2 probability: argmax parity is demanded only where this exceeds the tolerance."""
    t = prob.topk(2, dim=-1).values
    return t[..., 0] - t[..., 1]


# --------------------------------------------------------------------------------------
# style classifiers  (pitch_classifier.py:89-103, velocity_classifier.py:110-118, instrument_classifier.py:93-103)
# --------------------------------------------------------------------------------------
def classifier_forward(cfg: OracleConfig, p: Dict[str, Tensor], X: Tensor, num_layers: int = 2) -> Tensor:
    """Input(None, D) -> GRU x (num_layers - 1, return_sequences) -> GRU -> Dense(C, softmax): class probabilities (B, C)."""
    rnn, pre = (keras_gru, "gru") if cfg.cell_type == "GRU" else (keras_lstm, "lstm")
    h = X
    for k in range(1, num_layers):
        h = rnn(cfg, p, f"{pre}_{k}", h, True)
    h = rnn(cfg, p, f"{pre}_{num_layers}", h, False)
    return torch.softmax(h @ p["dense_1/kernel"] + p["dense_1/bias"], dim=-1)


def classifier_loss_and_grads(cfg: OracleConfig, p: Dict[str, Tensor], X: Tensor, Y: Tensor, num_layers: int = 2):
    """Keras categorical_crossentropy (renormalise, clip to [1e-7, 1 - 1e-7], mean over the batch) + accuracy; autograd gradients."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in p.items()}
    probs = classifier_forward(cfg, leaves, X, num_layers)
    q = probs / probs.sum(-1, keepdim=True)
    q = q.clamp(1e-7, 1 - 1e-7)
    loss = -(Y * q.log()).sum(-1).mean()
    acc = (probs.argmax(-1) == Y.argmax(-1)).double().mean()
    names = list(leaves)
    gl = torch.autograd.grad(loss, [leaves[k] for k in names])
    return {"loss": float(loss.detach()), "acc": float(acc)}, dict(zip(names, gl)), probs.detach()
