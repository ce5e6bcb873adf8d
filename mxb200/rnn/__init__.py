"""Legacy symbolic RNN API (reference: python/mxnet/rnn/{rnn_cell,rnn,io}.py)."""
from . import io, rnn, rnn_cell  # noqa: F401
from .io import *  # noqa: F401,F403
from .rnn import *  # noqa: F401,F403
from .rnn_cell import *  # noqa: F401,F403
