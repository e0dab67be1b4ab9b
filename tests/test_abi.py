"""CPU tests of the drop-in boundary: libmidivae.so builds for sm_100a, loads, exports every symbol that
include/midivae.h declares, and fails loudly (no fallback) when there is no B200."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "midivae.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mvae_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(lib_path):
    lib = C.CDLL(lib_path)
    declared = _declared_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/midivae.h but not exported"


def test_python_binding_covers_the_header(lib_path):
    from midi_vae_b200 import _lib
    assert sorted(_lib.SYMBOLS) == _declared_symbols()
    _lib.load()


def test_default_config_is_the_reference_settings(lib_path):
    from midi_vae_b200 import _lib
    lib = _lib.load()
    c = _lib.MvaeConfig()
    assert lib.mvae_default_config(C.byref(c)) == 0
    # settings.py:108-120,134,184,214 (LSTM branch of the reference defaults)
    assert (c.input_length, c.lstm_size, c.latent_rep_size, c.input_dim, c.meta_instrument_dim, c.meta_instrument_length) == (64, 256, 256, 61, 16, 4)
    assert (c.num_layers_encoder, c.num_layers_decoder, c.num_composers) == (2, 2, 2)
    assert abs(c.learning_rate - 2e-4) < 1e-7 and abs(c.beta - 0.1) < 1e-7 and abs(c.composer_weight - 0.1) < 1e-7
    assert abs(c.meta_instrument_weight - 0.1) < 1e-7 and c.meta_velocity_weight == 1.0
    assert c.gate_act == _lib.GATE["hard_sigmoid"]


def test_struct_layouts_match_header(lib_path):
    from midi_vae_b200 import _lib
    assert C.sizeof(_lib.MvaeConfig) == 18 * 4 + 11 * 4 + 3 * 4      # + cell_type, model_kind, cls_scalar_input
    assert C.sizeof(_lib.MvaeMetrics) == 40
    assert C.sizeof(_lib.MvaeBatch) == 8 + 8 * 8


def test_no_gpu_means_loud_failure(lib_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from midi_vae_b200 import Engine, EngineConfig, _lib
    with pytest.raises(_lib.MvaeError):
        Engine(EngineConfig(input_length=16, lstm_size=64, latent_rep_size=16, max_batch=4), 0)


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under midi_vae_b200/ may import or execute it."""
    pkg = os.path.join(ROOT, "midi_vae_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "midivae_oracle" not in src or f.endswith((".cu", ".cuh")) and "oracle/midivae_oracle.py" in src, f
