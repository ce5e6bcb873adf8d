"""Predefined models (reference: python/mxnet/gluon/model_zoo/)."""
from . import model_store  # noqa: F401
from . import vision  # noqa: F401
from .vision import get_model  # noqa: F401
