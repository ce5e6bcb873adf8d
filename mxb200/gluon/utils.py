"""Gluon utilities (reference: python/mxnet/gluon/utils.py)."""
import hashlib
import math
import os

import numpy as np
import torch

from ..ndarray import NDArray
from .. import ndarray as nd

__all__ = ["split_data", "split_and_load", "clip_global_norm", "check_sha1", "download"]


def _indent(s_, num_spaces):
    s = s_.split("\n")
    if len(s) == 1:
        return s_
    first = s.pop(0)
    s = [(num_spaces * " ") + line for line in s]
    return first + "\n" + "\n".join(s)


def split_data(data, num_slice, batch_axis=0, even_split=True):
    """Split along batch_axis into num_slice slices (utils.py:38-88)."""
    size = data.shape[batch_axis]
    if even_split and size % num_slice != 0:
        raise ValueError("data with shape %s cannot be evenly split into %d slices along axis %d. Use a batch size "
                         "that's multiple of %d or set even_split=False to allow uneven partitioning of data." % (
                             str(data.shape), num_slice, batch_axis, num_slice))
    if size < num_slice:
        raise ValueError("Too many slices for data with shape %s. Arguments are num_slice=%d and batch_axis=%d." % (
            str(data.shape), num_slice, batch_axis))
    step = size // num_slice
    if not even_split and size % num_slice != 0:
        # MXNet gives the remainder to the last slices; keep slices contiguous
        bounds = [int(round(i * size / float(num_slice))) for i in range(num_slice + 1)]
    else:
        bounds = [i * step for i in range(num_slice)] + [size]
    return [data.slice_axis(batch_axis, bounds[i], bounds[i + 1]) for i in range(num_slice)]


def split_and_load(data, ctx_list, batch_axis=0, even_split=True):
    """Split a batch and load each slice onto one context (utils.py:91-117)."""
    if not isinstance(data, NDArray):
        data = nd.array(data, ctx=ctx_list[0])
    if len(ctx_list) == 1:
        return [data.as_in_context(ctx_list[0])]
    slices = split_data(data, len(ctx_list), batch_axis, even_split)
    return [i.as_in_context(ctx) for i, ctx in zip(slices, ctx_list)]


def clip_global_norm(arrays, max_norm, check_isfinite=True):
    """Rescale arrays so that the sum of their 2-norms is at most max_norm (utils.py:120-150).  With
    ``check_isfinite=False`` no host synchronisation happens and the norm is returned as an NDArray."""
    assert len(arrays) > 0
    dev = arrays[0]._t.device
    total = torch.sqrt(sum((a._t.float() ** 2).sum().to(dev) for a in arrays))
    if not check_isfinite:
        scale = torch.clamp(max_norm / (total + 1e-8), max=1.0)
        from ..ndarray.sparse import BaseSparseNDArray
        with torch.no_grad():
            for a in arrays:
                if isinstance(a, BaseSparseNDArray):
                    a *= float(scale)
                else:
                    a._t.mul_(scale.to(a._t.device, a._t.dtype))
        from ..ndarray import NDArray
        return NDArray(total.reshape(1), arrays[0].context)
    total_f = float(total)
    if not np.isfinite(total_f):
        import warnings
        warnings.warn(UserWarning("nan or inf is detected. Clipping results will be undefined."), stacklevel=2)
    scale = max_norm / (total_f + 1e-8)
    if scale < 1.0:
        from ..ndarray.sparse import BaseSparseNDArray
        with torch.no_grad():
            for a in arrays:
                if isinstance(a, BaseSparseNDArray):
                    a *= scale                 # sparse storage: rescale the stored values
                else:
                    a._t.mul_(scale)
    return total_f


def check_sha1(filename, sha1_hash):
    sha1 = hashlib.sha1()
    with open(filename, "rb") as f:
        while True:
            data = f.read(1048576)
            if not data:
                break
            sha1.update(data)
    return sha1.hexdigest() == sha1_hash


def download(url, path=None, overwrite=False, sha1_hash=None, retries=5, verify_ssl=True):
    """No network in this environment: only returns an already-present local file."""
    fname = path if path and not os.path.isdir(path) else os.path.join(path or ".", url.split("/")[-1])
    if os.path.exists(fname) and not overwrite and (not sha1_hash or check_sha1(fname, sha1_hash)):
        return fname
    raise RuntimeError("download(%s): network access is not available; place the file at %s" % (url, fname))
