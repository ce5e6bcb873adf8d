"""ctypes bindings for the two native libraries (see build.py, csrc/c_api.h).

Loading rules (the driver records which in-tree .so each process loaded):
  * the CUDA library MUST load when a GPU is visible -- ops raise instead of silently
    falling back to eager torch on a GPU box;
  * on a CPU-only host the python fallbacks are used and the CUDA lib is optional.
"""
import ctypes
import os
import threading

import torch

from .base import MXNetError, getenv

_lock = threading.Lock()
_libs = {}
_HERE = os.path.dirname(os.path.abspath(__file__))

c_void_p, c_int, c_int64, c_float, c_char_p, c_size_t, c_double, c_uint64 = (
    ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_char_p, ctypes.c_size_t,
    ctypes.c_double, ctypes.c_uint64)


def _load(kind, required):
    with _lock:
        if kind in _libs:
            return _libs[kind]
        from . import build as B
        path = B.lib_path(kind)
        lib = None
        # MXB200_CUDA_LIB / MXB200_HOST_LIB: load an explicitly named build of the library (A/B runs of two kernel versions)
        override = os.environ.get("MXB200_%s_LIB" % kind.upper())
        try:
            if override:
                path = override
            elif not B.is_fresh(kind):
                if getenv("MXB200_NO_BUILD", False):
                    if not os.path.exists(path):
                        raise MXNetError("native library %s missing and MXB200_NO_BUILD set" % path)
                else:
                    B.build(kind)
            lib = ctypes.CDLL(path, mode=ctypes.RTLD_LOCAL)   # RTLD_GLOBAL breaks symbol binding next to libtorch
            lib.MXBGetLastError.restype = c_char_p
        except Exception as e:  # noqa: BLE001
            if required:
                raise MXNetError("failed to load native library '%s': %s" % (kind, e))
            lib = None
        _libs[kind] = lib
        return lib


def host(required=True):
    return _load("host", required)


def cuda(required=None):
    """The sm_100a kernel library.  Required whenever CUDA is available."""
    if required is None:
        required = torch.cuda.is_available()
    return _load("cuda", required)


def have_cuda_kernels():
    return torch.cuda.is_available() and cuda() is not None


_POST_CHECK = [None]


def check(lib, rc):
    if rc != 0:
        msg = lib.MXBGetLastError()
        raise MXNetError(msg.decode("utf-8", "replace") if msg else "native call failed (rc=%d)" % rc)
    if _POST_CHECK[0] is None:
        # MXB200_POST_KERNEL_CHECK=1: synchronise after every native launch so an asynchronous fault is reported at the
        # operator that caused it (the reference's MSHADOW_CUDA_POST_KERNEL_CHECK)
        _POST_CHECK[0] = bool(getenv("MXB200_POST_KERNEL_CHECK", False)) and torch.cuda.is_available()
    if _POST_CHECK[0] and lib is _libs.get("cuda"):
        torch.cuda.synchronize()


def ptr(t):
    """Device/host pointer of a tensor (or None) as c_void_p."""
    if t is None:
        return c_void_p(0)
    return c_void_p(t.data_ptr())


def stream_of(t=None, device=None):
    dev = t.device if t is not None else device
    return c_void_p(torch.cuda.current_stream(dev).cuda_stream)


class _ParamsIO(object):
    """Native .params writer/reader (csrc/params_io.cc)."""

    def __init__(self, lib):
        self.lib = lib

    def save(self, fname, names, tensors, ctxs):
        from .ndarray.serialization import TORCH2FLAG
        n = len(tensors)
        ptrs = (c_void_p * n)(*[t.data_ptr() for t in tensors])
        nbytes = (c_int64 * n)(*[t.numel() * t.element_size() for t in tensors])
        flags = (c_int * n)(*[TORCH2FLAG[t.dtype] for t in tensors])
        ndims = (c_int * n)(*[t.dim() for t in tensors])
        flat = [int(s) for t in tensors for s in t.shape]
        shapes = (c_int64 * max(1, len(flat)))(*flat)
        devt = (c_int * n)(*[c[0] for c in ctxs])
        devi = (c_int * n)(*[c[1] for c in ctxs])
        enc = [s.encode("utf-8") for s in names]
        cn = (c_char_p * max(1, len(enc)))(*enc)
        check(self.lib, self.lib.MXBParamsSave(fname.encode(), n, ptrs, nbytes, flags, ndims, shapes, devt, devi,
                                               len(enc), cn))


_pio = [False]


def params_io():
    if _pio[0] is False:
        lib = host(required=False)
        _pio[0] = _ParamsIO(lib) if lib is not None and hasattr(lib, "MXBParamsSave") else None
    return _pio[0]
