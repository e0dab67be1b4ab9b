"""Restatement of the slice of recurrentshop (github.com/farizrahman4u/recurrentshop, un-pinned master of early 2018 -- the
reference's README.md:38 installs it from git) that vae_definition.py:6-7,519-645 uses.  FROM MEMORY:
the original is not available offline.  TEST INFRASTRUCTURE -- see oracle/keras_shim/README.md.

``from recurrentshop import *`` also re-exports keras.layers (recurrentshop/engine.py does ``from keras.layers import *``),
which is where the reference's un-imported ``Activation`` (vae_definition.py:733) comes from."""
from keras.layers import *  # noqa: F401,F403
from keras.layers import Activation, Dense, Input, Lambda  # noqa: F401
from keras.models import Model  # noqa: F401

from .engine import RecurrentModel, RNNCell  # noqa: F401
from . import cells  # noqa: F401
from .cells import GRUCell, LSTMCell, SimpleRNNCell  # noqa: F401
