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


# ----------------------------------------------------------------------------- reference-compatible helpers (python/mxnet/base.py)
class NotImplementedForSymbol(MXNetError):
    """Raised by NDArray-only methods called on a Symbol (base.py:93-109)."""

    def __init__(self, function, alias, *args):
        super(NotImplementedForSymbol, self).__init__()
        self.function = function.__name__ if hasattr(function, "__name__") else str(function)
        self.alias = alias
        self.args = [str(type(a)) for a in args]

    def __str__(self):
        msg = "Function {}".format(self.function)
        if self.alias:
            msg += " (namely operator \"{}\")".format(self.alias)
        if self.args:
            msg += " with arguments ({})".format(", ".join(self.args))
        return msg + " is not implemented for Symbol and only available in NDArray."


class NotSupportedForSparseNDArray(MXNetError):
    """Raised by dense-only methods called on a sparse NDArray (base.py:111-127)."""

    def __init__(self, function, alias, *args):
        super(NotSupportedForSparseNDArray, self).__init__()
        self.function = function.__name__ if hasattr(function, "__name__") else str(function)
        self.alias = alias
        self.args = [str(type(a)) for a in args]

    def __str__(self):
        msg = "Function {}".format(self.function)
        if self.alias:
            msg += " (namely operator \"{}\")".format(self.alias)
        if self.args:
            msg += " with arguments ({})".format(", ".join(self.args))
        return msg + " is not supported for SparseNDArray and only available in NDArray."


class _NullType(object):
    """Placeholder for "argument not given" in generated operator signatures."""

    def __repr__(self):
        return "_Null"


_Null = _NullType()
mx_real_t = np.float32
mx_uint = int
mx_float = float


def py_str(x):
    return x.decode("utf-8") if isinstance(x, (bytes, bytearray)) else str(x)


def c_str(s):
    return s.encode("utf-8") if isinstance(s, str) else s


def data_dir_default():
    system = __import__("platform").system()
    if system == "Windows":
        return os.path.join(os.environ.get("APPDATA", os.path.expanduser("~")), "mxnet")
    return os.path.join(os.path.expanduser("~"), ".mxnet")


def data_dir():
    """Directory for datasets / pretrained models: $MXNET_HOME or ~/.mxnet (base.py:59-77)."""
    return os.getenv("MXNET_HOME", data_dir_default())


def _as_list(obj):
    return list(obj) if isinstance(obj, (list, tuple)) else [obj]


def check_call(ret):
    if ret != 0:
        raise MXNetError("native call failed (rc=%s)" % ret)


class classproperty(object):
    """Read-only property on the class itself (used by AttrScope / NameManager / Context ``_current``)."""

    def __init__(self, fget):
        self.fget = fget

    def __get__(self, obj, owner):
        return self.fget(owner)
