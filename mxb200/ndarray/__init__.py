"""``mx.nd`` namespace."""
import sys as _sys
import types as _types

from .ndarray import (NDArray, array, zeros, ones, empty, full, arange, zeros_like, ones_like, from_torch, waitall,
                      concatenate, moveaxis)
from . import op, sparse, random, serialization
from .serialization import save, load, load_frombuffer
from .sparse import RowSparseNDArray, CSRNDArray, BaseSparseNDArray, cast_storage
from .op import *  # noqa: F401,F403
from .op import _REG as _OPS

_this = _sys.modules[__name__]
for _k, _v in _OPS.items():
    if not hasattr(_this, _k) or _k in ("sum", "max", "min", "abs", "round", "slice"):
        setattr(_this, _k, _v)
# creation functions win over same-named elementwise ops
zeros_like = ndarray.zeros_like
ones_like = ndarray.ones_like

# mx.nd.contrib.* : ops registered as _contrib_X
contrib = _types.ModuleType(__name__ + ".contrib")
for _k, _v in _OPS.items():
    if _k.startswith("_contrib_"):
        setattr(contrib, _k[len("_contrib_"):], _v)
_sys.modules[contrib.__name__] = contrib

linalg = _types.ModuleType(__name__ + ".linalg")
for _k, _v in _OPS.items():
    if _k.startswith("linalg_"):
        setattr(linalg, _k[len("linalg_"):], _v)
_sys.modules[linalg.__name__] = linalg
