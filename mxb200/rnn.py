"""Legacy symbolic RNN cell API (reference: python/mxnet/rnn/) -- provided by the Gluon cells."""
from .gluon.rnn import *  # noqa: F401,F403
