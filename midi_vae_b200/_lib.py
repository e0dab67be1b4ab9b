"""ctypes binding of libmidivae.so (the C ABI declared in include/midivae.h).

There is no fallback: if the shared object is missing or cannot be loaded, importing a
symbol from here raises, and every compute entry point needs a B200 (sm_100a).
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmidivae.so")

NUM_METRICS = 10
GATE = {"hard_sigmoid": 0, "sigmoid": 1}
CELL = {"standard": 0, "recurrentshop_recalled": 1}
FEEDBACK = {"as_wired": 0, "teacher_forced": 1, "free_running": 2}
PRECISION = {"fp32": 0, "bf16": 1}
RNN_MODE = {"streamed": 0, "persistent": 1, "auto": 2}
CELL_TYPE = {"LSTM": 0, "GRU": 1}
PROF_CLASSES = ["rec_fwd", "rec_bwd", "gemm", "pointwise", "adam", "allreduce"]


class MvaeConfig(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "input_length", "lstm_size", "latent_rep_size", "input_dim", "meta_instrument_dim", "meta_instrument_length",
        "num_composers", "num_layers_encoder", "num_layers_decoder", "history", "extra_layer", "split_lstm_vector",
        "gate_act", "dec_cell_variant", "decoder_feedback", "precision", "rnn_mode", "max_batch")] + [(n, C.c_float) for n in (
        "beta", "prior_mean", "prior_std", "notes_weight", "meta_instrument_weight", "meta_velocity_weight", "composer_weight",
        "learning_rate", "adam_beta_1", "adam_beta_2", "adam_epsilon")] + [("cell_type", C.c_int), ("model_kind", C.c_int), ("cls_scalar_input", C.c_int)]


class MvaeBatch(C.Structure):
    _fields_ = [("n", C.c_int), ("pitch", C.c_void_p), ("target", C.c_void_p), ("instr", C.c_void_p), ("velocity", C.c_void_p),
                ("style", C.c_void_p), ("history", C.c_void_p), ("eps", C.c_void_p), ("w_notes", C.c_void_p)]


class MvaeMetrics(C.Structure):
    _fields_ = [("v", C.c_float * NUM_METRICS)]


class MvaeParamInfo(C.Structure):
    _fields_ = [("name", C.c_char * 96), ("offset", C.c_size_t), ("rows", C.c_int), ("cols", C.c_int), ("ld", C.c_int)]


# every symbol include/midivae.h declares: (restype, argtypes)
_H = C.c_void_p
_P = C.POINTER
SYMBOLS = {
    "mvae_version": (C.c_char_p, []),
    "mvae_default_config": (C.c_int, [_P(MvaeConfig)]),
    "mvae_create": (C.c_int, [_P(MvaeConfig), C.c_int, _P(_H)]),
    "mvae_destroy": (C.c_int, [_H]),
    "mvae_last_error": (C.c_char_p, [_H]),
    "mvae_param_tensor_count": (C.c_int, [_H, _P(C.c_int)]),
    "mvae_param_info_at": (C.c_int, [_H, C.c_int, _P(MvaeParamInfo)]),
    "mvae_arena_size": (C.c_int, [_H, _P(C.c_size_t)]),
    "mvae_get_param": (C.c_int, [_H, C.c_int, C.c_void_p]),
    "mvae_set_param": (C.c_int, [_H, C.c_int, C.c_void_p]),
    "mvae_commit_params": (C.c_int, [_H]),
    "mvae_get_grad": (C.c_int, [_H, C.c_int, C.c_void_p]),
    "mvae_reset_optimizer": (C.c_int, [_H]),
    "mvae_get_iterations": (C.c_int, [_H, _P(C.c_longlong)]),
    "mvae_train_step": (C.c_int, [_H, _P(MvaeBatch), C.c_void_p, C.c_void_p]),
    "mvae_train_step_host": (C.c_int, [_H, _P(MvaeBatch), _P(MvaeMetrics)]),
    "mvae_forward_backward": (C.c_int, [_H, _P(MvaeBatch), C.c_void_p, C.c_void_p]),
    "mvae_apply_update": (C.c_int, [_H, C.c_float, C.c_void_p]),
    "mvae_grad_arena": (C.c_int, [_H, _P(C.c_void_p), _P(C.c_size_t)]),
    "mvae_param_arena": (C.c_int, [_H, _P(C.c_void_p), _P(C.c_size_t)]),
    "mvae_eval_step": (C.c_int, [_H, _P(MvaeBatch), C.c_void_p, C.c_void_p]),
    "mvae_eval_step_host": (C.c_int, [_H, _P(MvaeBatch), _P(MvaeMetrics)]),
    "mvae_encode_host": (C.c_int, [_H, _P(MvaeBatch), C.c_void_p, C.c_void_p, C.c_void_p]),
    "mvae_decode_host": (C.c_int, [_H, _P(MvaeBatch), C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mvae_autoencode_host": (C.c_int, [_H, _P(MvaeBatch), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mvae_style_transfer": (C.c_int, [_H, _P(MvaeBatch), C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mvae_style_transfer_host": (C.c_int, [_H, _P(MvaeBatch), C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mvae_set_postprocess": (C.c_int, [_H, C.c_int, C.c_float, C.c_int, C.c_int]),
    "mvae_postprocess_host": (C.c_int, [_H, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "mvae_set_history_mode": (C.c_int, [_H, C.c_int]),
    "mvae_set_song_start_host": (C.c_int, [_H, C.c_void_p, C.c_int]),
    "mvae_cls_train_step_host": (C.c_int, [_H, _P(MvaeBatch), _P(MvaeMetrics)]),
    "mvae_cls_eval_step_host": (C.c_int, [_H, _P(MvaeBatch), _P(MvaeMetrics), C.c_void_p]),
    "mvae_nccl_unique_id": (C.c_int, [C.c_void_p]),
    "mvae_nccl_init": (C.c_int, [_H, C.c_void_p, C.c_int, C.c_int]),
    "mvae_world_size": (C.c_int, [_H, _P(C.c_int)]),
    "mvae_launch_count": (C.c_int, [_H, _P(C.c_longlong)]),
    "mvae_sync": (C.c_int, [_H]),
    "mvae_last_kernel_ms": (C.c_int, [_H, C.c_int, _P(C.c_float), _P(C.c_longlong)]),
    "mvae_set_profiling": (C.c_int, [_H, C.c_int]),
    "mvae_transfer_bytes": (C.c_int, [_H, _P(C.c_ulonglong), _P(C.c_ulonglong), C.c_int]),
    "mvae_stream": (C.c_int, [_H, _P(C.c_void_p)]),
    "mvae_selftest_gemm": (C.c_int, [C.c_int, C.c_int]),
}

_lib = None


def load() -> C.CDLL:
    """dlopen libmidivae.so and type every entry point.  Raises if the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m midi_vae_b200.build` (nvcc, sm_100a). "
            "midi_vae_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)      # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class MvaeError(RuntimeError):
    pass


def check(rc: int, handle=None) -> None:
    if rc != 0:
        msg = load().mvae_last_error(handle)
        raise MvaeError(f"libmidivae error {rc}: {msg.decode() if msg else '?'}")
