"""midi_vae_b200 -- B200-native MIDI-VAE hot path (train step + style-transfer inference).

Host-side mirror of the reference's model interface (``VAE.create`` -> ``encoder`` / ``decoder`` /
``autoencoder``) over the C ABI of libmidivae.so (include/midivae.h).  The CUDA library is the only
compute path: importing ``Engine``/``VAE`` works without a GPU, constructing them does not.
"""
from .engine import Engine, EngineConfig, METRIC_KEYS, reference_param_specs, nccl_unique_id  # noqa: F401
from .vae import VAE, History, initial_weights  # noqa: F401
from . import synth, marshal, postprocess, training  # noqa: F401

__version__ = "0.1.0"
