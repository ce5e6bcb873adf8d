"""Recurrent layers and cells (reference: python/mxnet/gluon/rnn/)."""
from .rnn_layer import *  # noqa: F401,F403
from .rnn_cell import *  # noqa: F401,F403
from .rnn_layer import _fused_rnn_forward  # noqa: F401
