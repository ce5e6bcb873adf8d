"""Shared base utilities: errors, dtype tables, env-flag access.

Parity notes: the reference funnels every C++ failure through ``MXNetError``
(python/mxnet/base.py) and reads configuration lazily with ``dmlc::GetEnv``
(docs/faq/env_var.md).  Here there is no C ABI; native pieces are torch C++
extensions, and ``getenv`` is the single lazily-read flag accessor.
"""
import os

import numpy as np
import torch

__all__ = ["MXNetError", "getenv", "string_types", "numeric_types", "integer_types",
           "bfloat16", "np_dtype", "torch_dtype", "dtype_name"]


class MXNetError(RuntimeError):
    """Error raised by the framework (name kept for reference-API parity)."""


string_types = (str,)
numeric_types = (float, int, np.generic)
integer_types = (int, np.integer)


class _BFloat16(object):
    """Sentinel dtype for bfloat16 (numpy has no native bf16)."""
    name = "bfloat16"
    itemsize = 2

    def __repr__(self):
        return "bfloat16"

    def __eq__(self, other):
        return other is self or other == "bfloat16" or other is torch.bfloat16

    def __hash__(self):
        return hash("bfloat16")


bfloat16 = _BFloat16()

_NP2TORCH = {
    np.dtype("float32"): torch.float32,
    np.dtype("float64"): torch.float64,
    np.dtype("float16"): torch.float16,
    np.dtype("uint8"): torch.uint8,
    np.dtype("int32"): torch.int32,
    np.dtype("int8"): torch.int8,
    np.dtype("int64"): torch.int64,
    np.dtype("bool"): torch.bool,
    np.dtype("int16"): torch.int16,
}
_TORCH2NP = {v: k for k, v in _NP2TORCH.items()}


def torch_dtype(dtype):
    """Anything dtype-like (str, numpy dtype, torch dtype, None) -> torch.dtype."""
    if dtype is None:
        return torch.float32
    if isinstance(dtype, torch.dtype):
        return dtype
    if dtype is bfloat16 or (isinstance(dtype, str) and dtype in ("bfloat16", "bf16")):
        return torch.bfloat16
    return _NP2TORCH[np.dtype(dtype)]


def np_dtype(dtype):
    """torch dtype (or dtype-like) -> numpy dtype; bf16 maps to the sentinel."""
    if isinstance(dtype, torch.dtype):
        if dtype is torch.bfloat16:
            return bfloat16
        return _TORCH2NP[dtype].type
    if dtype is bfloat16:
        return bfloat16
    return np.dtype(dtype).type


def dtype_name(dtype):
    t = torch_dtype(dtype)
    return "bfloat16" if t is torch.bfloat16 else _TORCH2NP[t].name


_TRUE = ("1", "true", "True", "TRUE", "yes", "on")


def getenv(name, default=None, typ=None):
    """Lazy env flag read (dmlc::GetEnv analogue).  ``typ`` inferred from default."""
    val = os.environ.get(name)
    if val is None:
        return default
    typ = typ or (type(default) if default is not None else str)
    if typ is bool:
        return val in _TRUE
    try:
        return typ(val)
    except ValueError:
        return default
