"""Restatement of the slice of Keras 2.0.8 (Theano backend semantics) that MIDI-VAE's vae_definition.py uses.
TEST INFRASTRUCTURE -- see oracle/keras_shim/README.md."""
__version__ = "2.0.8"          # the version pinned by the reference's shipped HDF5 checkpoints (restated, not the original)

from . import engine, initializers, backend, activations, objectives, optimizers, utils, layers, models  # noqa: F401,E402
from .layers import merge  # noqa: F401,E402
losses = objectives
