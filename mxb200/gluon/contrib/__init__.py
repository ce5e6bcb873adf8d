"""gluon.contrib (reference: python/mxnet/gluon/contrib/): experimental layers, RNN cells and samplers."""
from . import nn, rnn, data  # noqa: F401
