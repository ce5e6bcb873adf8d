"""Gluon: the imperative/hybrid high-level API (reference: python/mxnet/gluon/)."""
from .parameter import Parameter, Constant, ParameterDict, DeferredInitializationError  # noqa: F401
from .block import Block, HybridBlock, SymbolBlock  # noqa: F401
from . import nn, loss, utils  # noqa: F401
from .trainer import Trainer  # noqa: F401
from . import data, model_zoo, rnn, contrib  # noqa: F401
from ..parallel.graph import TrainStep  # noqa: F401,E402   (whole-step CUDA-graph capture)
