"""Engine: a thin numpy-facing wrapper over one libmidivae.so handle (one per GPU).

It owns no math.  Rolls go in PACKED (see synth.Rolls), results come back as numpy arrays.
Weight tensors are exchanged under the reference-style names of
``reference_param_specs`` (Keras layer names of the shipped checkpoints with gru -> lstm, and
descriptive names for the recurrentshop decoder cells); internally the 2*(nd+2) decoder
initial-state Denses are one column-fused tensor and this module maps both ways.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, asdict
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import _lib
from ._lib import MvaeBatch, MvaeConfig, MvaeMetrics, MvaeParamInfo, check

METRIC_KEYS = ["loss", "decoder_loss_1", "decoder_loss_2", "decoder_loss_3", "composer_decoder_loss",
               "decoder_acc_1", "decoder_acc_2", "decoder_acc_3", "composer_decoder_acc", "kl"]


@dataclass
class EngineConfig:
    """Mirrors mvae_config (include/midivae.h); defaults = settings.py with the LSTM branch."""
    input_length: int = 64
    lstm_size: int = 256
    latent_rep_size: int = 256
    input_dim: int = 61
    meta_instrument_dim: int = 16
    meta_instrument_length: int = 4
    num_composers: int = 2
    num_layers_encoder: int = 2
    num_layers_decoder: int = 2
    history: bool = True
    extra_layer: bool = True
    split_lstm_vector: bool = True
    gate_act: str = "hard_sigmoid"
    dec_cell_variant: str = "recurrentshop_recalled"   # what the reference decoder computes (recurrentshop LSTMCell); "standard" = Keras equations, an extension
    decoder_feedback: str = "as_wired"
    precision: str = "fp32"
    rnn_mode: str = "auto"
    max_batch: int = 256
    beta: float = 0.1
    prior_mean: float = 0.0
    prior_std: float = 1.0
    notes_weight: float = 1.0
    meta_instrument_weight: float = 0.1
    meta_velocity_weight: float = 1.0
    composer_weight: float = 0.1
    learning_rate: float = 2e-4
    adam_beta_1: float = 0.9
    adam_beta_2: float = 0.999
    adam_epsilon: float = 1e-8
    cell_type: str = "LSTM"        # 'LSTM' (north_star; cluster kernels) | 'GRU' (the reference's shipped default, settings.py:155)

    def to_c(self) -> MvaeConfig:
        c = MvaeConfig()
        for k, v in asdict(self).items():
            if k == "gate_act": v = _lib.GATE[v]
            elif k == "dec_cell_variant": v = _lib.CELL[v]
            elif k == "decoder_feedback": v = _lib.FEEDBACK[v]
            elif k == "precision": v = _lib.PRECISION[v]
            elif k == "rnn_mode": v = _lib.RNN_MODE[v]
            elif k == "cell_type": v = _lib.CELL_TYPE[v]
            setattr(c, k, int(v) if isinstance(v, bool) else v)
        return c


def reference_param_specs(cfg: EngineConfig) -> List[Tuple[str, Tuple[int, ...]]]:
    """(name, shape) in the reference checkpoint order: encoder layers, then the decoder's
    initial-state Denses (notes L1.., instrument, velocity), notes cells + Dense, instrument cell + Dense,
    velocity cell + Dense.  Per recurrentshop cell: kernel, bias, recurrent_kernel."""
    gru = cfg.cell_type == "GRU"
    H, L, G = cfg.lstm_size, cfg.latent_rep_size, (3 if gru else 4) * cfg.lstm_size
    Dp, Di = cfg.input_dim, cfg.meta_instrument_dim
    pre = "gru" if gru else "lstm"
    s: List[Tuple[str, Tuple[int, ...]]] = []

    def keras_lstm(name, D):
        s.extend([(f"{name}/kernel", (D, G)), (f"{name}/recurrent_kernel", (H, G)), (f"{name}/bias", (G,))])
    for k in range(1, cfg.num_layers_encoder + 1):
        keras_lstm(f"{pre}_{k}", Dp if k == 1 else H)
    keras_lstm(f"{pre}_meta_instrument", Di)
    keras_lstm(f"{pre}_meta_velocity", 1)
    s.extend([("extra_instrument_after_concat_layer/kernel", (3 * H, H)), ("extra_instrument_after_concat_layer/bias", (H,))])
    if cfg.extra_layer:
        s.extend([("extra_layer/kernel", (H, H)), ("extra_layer/bias", (H,))])
    half = H // 2
    s.extend([("z_mean/kernel", (half, L)), ("z_mean/bias", (L,)), ("z_log_var/kernel", (H - half, L)), ("z_log_var/bias", (L,))])
    Q = 2 * L if cfg.history else L
    for nm in [f"notes_l{k}" for k in range(1, cfg.num_layers_decoder + 1)] + ["instr", "vel"]:
        for j in ((1,) if gru else (1, 2)):          # one initial-state Dense per cell state (vae_definition.py:563-568)
            s.extend([(f"dec_init/{nm}_s{j}/kernel", (Q, H)), (f"dec_init/{nm}_s{j}/bias", (H,))])

    def rs_cell(name, D):
        s.extend([(f"{name}/kernel", (D, G)), (f"{name}/bias", (G,))])
        if gru:      # recurrentshop GRUCell: Dense(2H) on h for z, r and Dense(H) on r*h (the shipped checkpoints' dense_k+1, dense_k+2)
            s.extend([(f"{name}/recurrent_kernel_1", (H, 2 * H)), (f"{name}/recurrent_kernel_2", (H, H))])
        else:
            s.extend([(f"{name}/recurrent_kernel", (H, G))])
    for k in range(1, cfg.num_layers_decoder + 1):
        rs_cell(f"notes/cell_{k}", Dp if k == 1 else H)
    s.extend([("notes/out/kernel", (H, Dp)), ("notes/out/bias", (Dp,))])
    rs_cell("meta_instrument/cell", Di)
    s.extend([("meta_instrument/out/kernel", (H, Di)), ("meta_instrument/out/bias", (Di,))])
    rs_cell("meta_velocity/cell", 1)
    s.extend([("meta_velocity/out/kernel", (H, 1)), ("meta_velocity/out/bias", (1,))])
    return s


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _u8(a, shape) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint8)
    if a.shape != tuple(shape):
        raise ValueError(f"expected uint8 array of shape {tuple(shape)}, got {a.shape}")
    return a


def _f32(a, shape) -> Optional[np.ndarray]:
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float32)
    if a.shape != tuple(shape):
        raise ValueError(f"expected float32 array of shape {tuple(shape)}, got {a.shape}")
    return a


class Engine:
    def __init__(self, cfg: EngineConfig, device: int = 0):
        self.cfg = cfg
        self.device = device
        self.lib = _lib.load()
        self._h = C.c_void_p()
        c = cfg.to_c()
        rc = self.lib.mvae_create(C.byref(c), device, C.byref(self._h))
        if rc != 0:
            msg = self.lib.mvae_last_error(None)
            self._h = None
            raise _lib.MvaeError(f"mvae_create failed ({rc}): {msg.decode() if msg else '?'}")
        n = C.c_int()
        check(self.lib.mvae_param_tensor_count(self._h, C.byref(n)), self._h)
        self._table: List[Tuple[str, int, int]] = []
        for i in range(n.value):
            info = MvaeParamInfo()
            check(self.lib.mvae_param_info_at(self._h, i, C.byref(info)), self._h)
            self._table.append((info.name.decode(), info.rows, info.cols))
        self._index = {nm: i for i, (nm, _, _) in enumerate(self._table)}
        self._keep = []   # arrays referenced by the batch struct of the call in flight

    # ------------------------------------------------------------------ lifecycle
    def close(self):
        if getattr(self, "_h", None):
            self.lib.mvae_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    # ------------------------------------------------------------------ weights
    def internal_table(self):
        return list(self._table)

    def _get_internal(self, name, grad=False) -> np.ndarray:
        i = self._index[name]
        _, r, c = self._table[i]
        out = np.empty((r, c), np.float32)
        fn = self.lib.mvae_get_grad if grad else self.lib.mvae_get_param
        check(fn(self._h, i, _ptr(out)), self._h)
        return out

    def _set_internal(self, name, arr):
        i = self._index[name]
        _, r, c = self._table[i]
        a = np.ascontiguousarray(arr, np.float32).reshape(r, c)
        check(self.lib.mvae_set_param(self._h, i, _ptr(a)), self._h)

    def _init_names(self):
        return [f"notes_l{k}" for k in range(1, self.cfg.num_layers_decoder + 1)] + ["instr", "vel"]

    def _init_cols(self):
        """dec_init/<cell>_s<j> -> first column of that Dense inside the column-fused dec_init tensors."""
        H = self.cfg.lstm_size
        spc = 1 if self.cfg.cell_type == "GRU" else 2           # states per cell
        return {f"dec_init/{nm}_s{j}": (spc * r + (j - 1)) * H for r, nm in enumerate(self._init_names()) for j in range(1, spc + 1)}

    def get_weights(self, grad: bool = False) -> Dict[str, np.ndarray]:
        """Reference-named tensors (or their gradients from the last forward_backward)."""
        H = self.cfg.lstm_size
        out: Dict[str, np.ndarray] = {}
        fused_k = self._get_internal("dec_init/kernel", grad)
        fused_b = self._get_internal("dec_init/bias", grad)
        col = self._init_cols()
        cache: Dict[str, np.ndarray] = {}
        for name, shape in reference_param_specs(self.cfg):
            if name.startswith("dec_init/"):
                base, kind = name.rsplit("/", 1)
                c0 = col[base]
                out[name] = (fused_k[:, c0:c0 + H] if kind == "kernel" else fused_b[0, c0:c0 + H]).copy()
            elif name.endswith(("/recurrent_kernel_1", "/recurrent_kernel_2")):
                # recurrentshop GRUCell: the (H,2H) z,r kernel and the (H,H) candidate kernel are one (H,3H) tensor [z|r|h] in the arena
                base = name.rsplit("/", 1)[0] + "/recurrent_kernel"
                if base not in cache:
                    cache[base] = self._get_internal(base, grad)
                out[name] = (cache[base][:, :2 * H] if name.endswith("_1") else cache[base][:, 2 * H:]).copy()
            else:
                out[name] = self._get_internal(name, grad).reshape(shape).copy()
        return out

    def set_weights(self, weights: Dict[str, np.ndarray]) -> None:
        H = self.cfg.lstm_size
        specs = reference_param_specs(self.cfg)
        missing = [n for n, _ in specs if n not in weights]
        if missing:
            raise KeyError(f"missing weight tensors: {missing[:4]}...")
        Q = 2 * self.cfg.latent_rep_size if self.cfg.history else self.cfg.latent_rep_size
        col = self._init_cols()
        nS = len(col)
        fused_k = np.zeros((Q, nS * H), np.float32)
        fused_b = np.zeros((1, nS * H), np.float32)
        fused_u: Dict[str, np.ndarray] = {}
        for name, shape in specs:
            w = np.asarray(weights[name], np.float32)
            if w.shape != tuple(shape):
                raise ValueError(f"{name}: expected shape {shape}, got {w.shape}")
            if name.startswith("dec_init/"):
                base, kind = name.rsplit("/", 1)
                c0 = col[base]
                if kind == "kernel":
                    fused_k[:, c0:c0 + H] = w
                else:
                    fused_b[0, c0:c0 + H] = w
            elif name.endswith(("/recurrent_kernel_1", "/recurrent_kernel_2")):
                u = fused_u.setdefault(name.rsplit("/", 1)[0] + "/recurrent_kernel", np.zeros((H, 3 * H), np.float32))
                if name.endswith("_1"):
                    u[:, :2 * H] = w
                else:
                    u[:, 2 * H:] = w
            else:
                self._set_internal(name, w)
        for name, u in fused_u.items():
            self._set_internal(name, u)
        self._set_internal("dec_init/kernel", fused_k)
        self._set_internal("dec_init/bias", fused_b)
        check(self.lib.mvae_commit_params(self._h), self._h)

    def get_grads(self) -> Dict[str, np.ndarray]:
        return self.get_weights(grad=True)

    def reset_optimizer(self):
        check(self.lib.mvae_reset_optimizer(self._h), self._h)

    @property
    def iterations(self) -> int:
        t = C.c_longlong()
        check(self.lib.mvae_get_iterations(self._h, C.byref(t)), self._h)
        return t.value

    # ------------------------------------------------------------------ batches
    def _batch(self, pitch, instr, velocity, style=None, history=None, eps=None, w_notes=None, target=None) -> MvaeBatch:
        cfg = self.cfg
        n = int(np.asarray(pitch).shape[0])
        T, Ti, L = cfg.input_length, cfg.meta_instrument_length, cfg.latent_rep_size
        arrs = dict(
            pitch=_u8(pitch, (n, T)), target=None if target is None else _u8(target, (n, T)), instr=_u8(instr, (n, Ti)),
            velocity=_f32(velocity, (n, T)), style=None if style is None else _u8(style, (n,)),
            history=_f32(history, (n, L)), eps=_f32(eps, (n, L)), w_notes=_f32(w_notes, (n, T)))
        for nm, hi in (("pitch", cfg.input_dim), ("target", cfg.input_dim), ("instr", cfg.meta_instrument_dim), ("style", cfg.num_composers)):
            if arrs[nm] is not None and arrs[nm].size and int(arrs[nm].max()) >= hi:
                raise ValueError(f"{nm} class index out of range (max {int(arrs[nm].max())} >= {hi})")
        self._keep = list(arrs.values())
        b = MvaeBatch()
        b.n = n
        for k, v in arrs.items():
            setattr(b, k, None if v is None else v.ctypes.data)
        return b

    def _metrics(self, m: MvaeMetrics) -> Dict[str, float]:
        return {k: float(m.v[i]) for i, k in enumerate(METRIC_KEYS)}

    # ------------------------------------------------------------------ hot path (host-pointer API)
    def set_history_mode(self, from_batch: bool):
        """Opt-in: True = train / evaluate steps take the decoder's history latents from the step's own z (H[i] = z[i-1] inside a
        song; pass ``song_start`` to the step) instead of the ``history`` argument the reference fills from a separate encoder pass."""
        check(self.lib.mvae_set_history_mode(self._h, 1 if from_batch else 0), self._h)
        self._history_from_batch = bool(from_batch)

    def _song_start(self, song_start, n):
        if not getattr(self, "_history_from_batch", False):
            if song_start is not None:
                raise ValueError("song_start is only used with set_history_mode(True)")
            return
        ss = None if song_start is None else _u8(np.asarray(song_start).astype(np.uint8), (n,))
        check(self.lib.mvae_set_song_start_host(self._h, _ptr(ss), n if ss is not None else 0), self._h)

    def train_on_batch(self, pitch, instr, velocity, style, history=None, eps=None, w_notes=None, target=None, song_start=None) -> Dict[str, float]:
        b = self._batch(pitch, instr, velocity, style, history, eps, w_notes, target)
        self._song_start(song_start, b.n)
        m = MvaeMetrics()
        check(self.lib.mvae_train_step_host(self._h, C.byref(b), C.byref(m)), self._h)
        return self._metrics(m)

    def evaluate_batch(self, pitch, instr, velocity, style, history=None, eps=None, w_notes=None, target=None, song_start=None) -> Dict[str, float]:
        b = self._batch(pitch, instr, velocity, style, history, eps, w_notes, target)
        self._song_start(song_start, b.n)
        m = MvaeMetrics()
        check(self.lib.mvae_eval_step_host(self._h, C.byref(b), C.byref(m)), self._h)
        return self._metrics(m)

    def encode(self, pitch, instr, velocity, eps=None):
        b = self._batch(pitch, instr, velocity, eps=eps)
        L = self.cfg.latent_rep_size
        z, mu, lv = (np.empty((b.n, L), np.float32) for _ in range(3))
        check(self.lib.mvae_encode_host(self._h, C.byref(b), _ptr(z), _ptr(mu), _ptr(lv)), self._h)
        return z, mu, lv

    def decode(self, z, history=None, feedback: str = "as_wired", pitch=None, instr=None, velocity=None):
        cfg = self.cfg
        z = np.ascontiguousarray(z, np.float32)
        n = z.shape[0]
        if z.shape != (n, cfg.latent_rep_size):
            raise ValueError("z must be (n, latent_rep_size)")
        history = _f32(history, (n, cfg.latent_rep_size))
        if feedback == "teacher_forced":
            b = self._batch(pitch, instr, velocity)
        else:
            b = MvaeBatch(); b.n = n
        Y = np.empty((n, cfg.input_length, cfg.input_dim), np.float32)
        I = np.empty((n, cfg.meta_instrument_length, cfg.meta_instrument_dim), np.float32)
        V = np.empty((n, cfg.input_length), np.float32)
        check(self.lib.mvae_decode_host(self._h, C.byref(b), _ptr(z), _ptr(history), _lib.FEEDBACK[feedback], _ptr(Y), _ptr(I), _ptr(V)), self._h)
        return Y, I, V

    def autoencode(self, pitch, instr, velocity, history=None, eps=None, target=None):
        cfg = self.cfg
        b = self._batch(pitch, instr, velocity, None, history, eps, None, target)
        n = b.n
        Y = np.empty((n, cfg.input_length, cfg.input_dim), np.float32)
        I = np.empty((n, cfg.meta_instrument_length, cfg.meta_instrument_dim), np.float32)
        V = np.empty((n, cfg.input_length), np.float32)
        S = np.empty((n, cfg.num_composers), np.float32)
        z = np.empty((n, cfg.latent_rep_size), np.float32)
        check(self.lib.mvae_autoencode_host(self._h, C.byref(b), _ptr(Y), _ptr(I), _ptr(V), _ptr(S), _ptr(z)), self._h)
        return Y, I, V, S, z

    _SCOPES = {None: 0, "off": 0, "chunk": 1, "song": 2}

    def set_postprocess(self, scope=None, velocity_threshold: float = 0.5, override_by_velocity: bool = True, max_voices: int = 4):
        """Device-side process_decoder_outputs rules (vae_definition.py:1156-1190) for every following style_transfer call: scope 'chunk'
        (the voice memory restarts per chunk, as in the reference's style-switch loop), 'song' (per song) or None (raw argmax / velocities)."""
        check(self.lib.mvae_set_postprocess(self._h, self._SCOPES[scope], velocity_threshold, int(override_by_velocity), max_voices), self._h)

    def postprocess(self, pitch, velocity, song_start=None, scope: str = "chunk"):
        """The same rules on packed rolls supplied by the caller: returns (velocity', held) with held = 1 where no note is struck."""
        cfg = self.cfg
        P = np.ascontiguousarray(np.asarray(pitch, dtype=np.uint8).reshape(-1, cfg.input_length))
        n = P.shape[0]
        V = np.ascontiguousarray(np.asarray(velocity, np.float32).reshape(n, cfg.input_length)).copy()
        ss = None if song_start is None else _u8(np.asarray(song_start).astype(np.uint8), (n,))
        D = np.empty((n, cfg.input_length), np.uint8)
        check(self.lib.mvae_postprocess_host(self._h, n, _ptr(P), _ptr(ss), self._SCOPES[scope], _ptr(V), _ptr(D)), self._h)
        return V, D

    def style_transfer(self, pitch, instr, velocity, c_from=0, c_to=1, song_start=None, feedback: str = "as_wired"):
        cfg = self.cfg
        b = self._batch(pitch, instr, velocity)
        n = b.n
        ss = None if song_start is None else _u8(np.asarray(song_start).astype(np.uint8), (n,))
        P = np.empty((n, cfg.input_length), np.uint8)
        I = np.empty((n, cfg.meta_instrument_length), np.uint8)
        V = np.empty((n, cfg.input_length), np.float32)
        check(self.lib.mvae_style_transfer_host(self._h, C.byref(b), _ptr(ss), c_from, c_to, _lib.FEEDBACK[feedback], _ptr(P), _ptr(I), _ptr(V)), self._h)
        return P, I, V

    # ------------------------------------------------------------------ device-pointer API (bench, DP)
    def device_batch(self, n, pitch, instr, velocity, style=None, history=None, eps=None, w_notes=None, target=None) -> MvaeBatch:
        """All arguments are raw device addresses (ints) or None."""
        b = MvaeBatch()
        b.n = n
        b.pitch, b.instr, b.velocity, b.style, b.history, b.eps, b.w_notes, b.target = pitch, instr, velocity, style, history, eps, w_notes, target
        return b

    def train_step_device(self, b: MvaeBatch, metrics_ptr=None, stream=None):
        check(self.lib.mvae_train_step(self._h, C.byref(b), metrics_ptr, stream), self._h)

    def forward_backward_device(self, b: MvaeBatch, metrics_ptr=None, stream=None):
        check(self.lib.mvae_forward_backward(self._h, C.byref(b), metrics_ptr, stream), self._h)

    def apply_update(self, grad_scale: float = 1.0, stream=None):
        check(self.lib.mvae_apply_update(self._h, grad_scale, stream), self._h)

    def eval_step_device(self, b: MvaeBatch, metrics_ptr=None, stream=None):
        check(self.lib.mvae_eval_step(self._h, C.byref(b), metrics_ptr, stream), self._h)

    def style_transfer_device(self, b: MvaeBatch, song_start_ptr, c_from, c_to, feedback, pitch_ptr, instr_ptr, vel_ptr, stream=None):
        check(self.lib.mvae_style_transfer(self._h, C.byref(b), song_start_ptr, c_from, c_to, _lib.FEEDBACK[feedback], pitch_ptr, instr_ptr, vel_ptr, stream), self._h)

    def grad_arena(self) -> Tuple[int, int]:
        p, n = C.c_void_p(), C.c_size_t()
        check(self.lib.mvae_grad_arena(self._h, C.byref(p), C.byref(n)), self._h)
        return p.value, n.value

    def stream(self) -> int:
        s = C.c_void_p()
        check(self.lib.mvae_stream(self._h, C.byref(s)), self._h)
        return s.value

    def sync(self):
        check(self.lib.mvae_sync(self._h), self._h)

    def launch_count(self) -> int:
        n = C.c_longlong()
        check(self.lib.mvae_launch_count(self._h, C.byref(n)), self._h)
        return n.value

    def set_profiling(self, on: bool):
        check(self.lib.mvae_set_profiling(self._h, int(on)), self._h)

    def kernel_ms(self) -> Dict[str, Tuple[float, int]]:
        out = {}
        for i, name in enumerate(_lib.PROF_CLASSES):
            ms, n = C.c_float(), C.c_longlong()
            check(self.lib.mvae_last_kernel_ms(self._h, i, C.byref(ms), C.byref(n)), self._h)
            out[name] = (ms.value, n.value)
        return out

    def transfer_bytes(self, reset=False) -> Tuple[int, int]:
        a, b = C.c_ulonglong(), C.c_ulonglong()
        check(self.lib.mvae_transfer_bytes(self._h, C.byref(a), C.byref(b), int(reset)), self._h)
        return a.value, b.value

    # ------------------------------------------------------------------ data parallel
    def nccl_init(self, unique_id: bytes, world_size: int, rank: int):
        buf = C.create_string_buffer(unique_id, 128)
        check(self.lib.mvae_nccl_init(self._h, buf, world_size, rank), self._h)


def nccl_unique_id() -> bytes:
    lib = _lib.load()
    buf = C.create_string_buffer(128)
    check(lib.mvae_nccl_unique_id(buf), None)
    return buf.raw
