""".params / NDArray list serialization, byte-compatible with the reference.

Format (reference: src/ndarray/ndarray.cc:1574-1808), little endian:

  file   := u64 0x112 | u64 0 | u64 n | n * blob | u64 m | m * (u64 len | bytes)
  blobV2 := u32 0xF993fac9 | i32 stype | [sparse: storage_shape] | shape
            | (stop if ndim == 0) | i32 dev_type | i32 dev_id | i32 type_flag
            | [sparse: per aux (i32 type | shape)] | data bytes | [sparse: aux bytes]
  shape  := u32 ndim | ndim * i64

The loader also accepts V1 (0xF993fac8) and the pre-magic legacy layout where
the first u32 is ndim followed by u32 dims.  bf16 has no flag in MXNet 1.3.1;
we use 12 (what later upstream releases chose) and offer ``dtype=`` down/up
casting on save for files the original runtime must read.

A C++ fast path (mxb200/csrc/params_io.cc) implements the same format for
large checkpoints; this module is the portable implementation and the oracle
for its tests.
"""
import io
import struct

import numpy as np
import torch

from ..base import MXNetError
from ..context import Context
from .ndarray import NDArray

LIST_MAGIC = 0x112
V1_MAGIC = 0xF993FAC8
V2_MAGIC = 0xF993FAC9

# mshadow type flags (3rdparty/mshadow/mshadow/base.h:303-309) + bf16 = 12
FLAG2TORCH = {0: torch.float32, 1: torch.float64, 2: torch.float16, 3: torch.uint8, 4: torch.int32,
              5: torch.int8, 6: torch.int64, 12: torch.bfloat16}
TORCH2FLAG = {v: k for k, v in FLAG2TORCH.items()}
TORCH2FLAG[torch.bool] = 3
STYPE = {"default": 0, "row_sparse": 1, "csr": 2}


def _write_shape(f, shape):
    f.write(struct.pack("<I", len(shape)))
    if len(shape):
        f.write(struct.pack("<%dq" % len(shape), *shape))


def _read_exact(f, n):
    b = f.read(n)
    if len(b) != n:
        raise MXNetError("Invalid NDArray file format (truncated)")
    return b


def _read_shape(f):
    ndim, = struct.unpack("<I", _read_exact(f, 4))
    return tuple(struct.unpack("<%dq" % ndim, _read_exact(f, 8 * ndim))) if ndim else ()


def _raw_bytes(t):
    t = t.detach().contiguous().cpu()
    if t.dtype is torch.bfloat16:
        return t.view(torch.int16).numpy().tobytes()
    if t.dtype is torch.bool:
        t = t.to(torch.uint8)
    return t.numpy().tobytes()


def _from_bytes(buf, flag, shape):
    td = FLAG2TORCH.get(flag)
    if td is None:
        raise MXNetError("unknown type flag %d" % flag)
    n = int(np.prod(shape)) if len(shape) else 1
    if td is torch.bfloat16:
        a = np.frombuffer(buf, dtype=np.int16, count=n).copy()
        return torch.from_numpy(a).view(torch.bfloat16).reshape(shape)
    npd = {torch.float32: np.float32, torch.float64: np.float64, torch.float16: np.float16, torch.uint8: np.uint8,
           torch.int32: np.int32, torch.int8: np.int8, torch.int64: np.int64}[td]
    a = np.frombuffer(buf, dtype=npd, count=n).copy()
    return torch.from_numpy(a).reshape(shape)


def _itemsize(flag):
    return {0: 4, 1: 8, 2: 2, 3: 1, 4: 4, 5: 1, 6: 8, 12: 2}[flag]


def save_ndarray(f, arr, dtype=None):
    from . import sparse
    f.write(struct.pack("<Ii", V2_MAGIC, STYPE[arr.stype]))
    ctx = arr.context
    if isinstance(arr, sparse.RowSparseNDArray):
        data = arr._data if dtype is None else arr._data.to(dtype)
        _write_shape(f, tuple(data.shape))
        _write_shape(f, arr.shape)
        f.write(struct.pack("<iii", ctx.device_typeid, ctx.device_id, TORCH2FLAG[data.dtype]))
        f.write(struct.pack("<i", 6)); _write_shape(f, tuple(arr._indices.shape))
        f.write(_raw_bytes(data)); f.write(_raw_bytes(arr._indices))
        return
    if isinstance(arr, sparse.CSRNDArray):
        data = arr._data if dtype is None else arr._data.to(dtype)
        _write_shape(f, tuple(data.shape))
        _write_shape(f, arr.shape)
        f.write(struct.pack("<iii", ctx.device_typeid, ctx.device_id, TORCH2FLAG[data.dtype]))
        f.write(struct.pack("<i", 6)); _write_shape(f, tuple(arr._indptr.shape))
        f.write(struct.pack("<i", 6)); _write_shape(f, tuple(arr._indices.shape))
        f.write(_raw_bytes(data)); f.write(_raw_bytes(arr._indptr)); f.write(_raw_bytes(arr._indices))
        return
    t = arr._t if dtype is None else arr._t.to(dtype)
    _write_shape(f, tuple(t.shape))
    if t.dim() == 0:
        return
    f.write(struct.pack("<iii", ctx.device_typeid, ctx.device_id, TORCH2FLAG[t.dtype]))
    f.write(_raw_bytes(t))


def _load_legacy(f, magic, ctx):
    if magic == V1_MAGIC:
        shape = _read_shape(f)
    else:
        ndim = magic
        shape = tuple(struct.unpack("<%dI" % ndim, _read_exact(f, 4 * ndim))) if ndim else ()
    if len(shape) == 0:
        return None
    _dt, _di, flag = struct.unpack("<iii", _read_exact(f, 12))
    n = int(np.prod(shape))
    t = _from_bytes(_read_exact(f, n * _itemsize(flag)), flag, shape)
    return NDArray(t.to(ctx.torch_device), ctx)


def load_ndarray(f, ctx=None):
    from . import sparse
    ctx = ctx or Context("cpu", 0)
    magic, = struct.unpack("<I", _read_exact(f, 4))
    if magic != V2_MAGIC:
        return _load_legacy(f, magic, ctx)
    stype, = struct.unpack("<i", _read_exact(f, 4))
    nad = {0: 0, 1: 1, 2: 2}[stype]
    sshape = _read_shape(f) if nad else None
    shape = _read_shape(f)
    if len(shape) == 0:
        return None
    _dt, _di, flag = struct.unpack("<iii", _read_exact(f, 12))
    aux = []
    for _ in range(nad):
        at, = struct.unpack("<i", _read_exact(f, 4))
        aux.append((at, _read_shape(f)))
    dshape = sshape if nad else shape
    n = int(np.prod(dshape)) if len(dshape) else 0
    data = _from_bytes(_read_exact(f, n * _itemsize(flag)), flag, dshape) if n else torch.zeros(dshape, dtype=FLAG2TORCH[flag])
    auxd = []
    for at, ash in aux:
        an = int(np.prod(ash)) if len(ash) else 0
        auxd.append(_from_bytes(_read_exact(f, an * _itemsize(at)), at, ash) if an else torch.zeros(ash, dtype=torch.int64))
    d = ctx.torch_device
    if stype == 0:
        return NDArray(data.to(d), ctx)
    if stype == 1:
        return sparse.RowSparseNDArray(data.to(d), auxd[0].to(d), shape, ctx)
    return sparse.CSRNDArray(data.to(d), auxd[1].to(d), auxd[0].to(d), shape, ctx)


def dumps(data, dtype=None):
    """Serialize a list / dict of NDArrays to bytes."""
    if isinstance(data, NDArray) or hasattr(data, "stype") and not isinstance(data, (dict, list, tuple)):
        data = [data]
    if isinstance(data, dict):
        names, arrs = list(data.keys()), list(data.values())
        for k in names:
            if not isinstance(k, str):
                raise TypeError("save only accept dict str->NDArray or list of NDArray")
    else:
        names, arrs = [], list(data)
    for a in arrs:
        if not hasattr(a, "stype"):
            raise TypeError("save only accept dict str->NDArray or list of NDArray")
    td = None
    if dtype is not None:
        from ..base import torch_dtype
        td = torch_dtype(dtype)
    f = io.BytesIO()
    f.write(struct.pack("<QQQ", LIST_MAGIC, 0, len(arrs)))
    for a in arrs:
        save_ndarray(f, a, td if (td is not None and a.stype == "default" and a._t.is_floating_point()) else None)
    f.write(struct.pack("<Q", len(names)))
    for k in names:
        b = k.encode("utf-8")
        f.write(struct.pack("<Q", len(b)))
        f.write(b)
    return f.getvalue()


def loads(buf, ctx=None):
    f = io.BytesIO(buf)
    header, _res = struct.unpack("<QQ", _read_exact(f, 16))
    if header != LIST_MAGIC:
        raise MXNetError("Invalid NDArray file format")
    n, = struct.unpack("<Q", _read_exact(f, 8))
    arrs = [load_ndarray(f, ctx) for _ in range(n)]
    m, = struct.unpack("<Q", _read_exact(f, 8))
    names = []
    for _ in range(m):
        ln, = struct.unpack("<Q", _read_exact(f, 8))
        names.append(_read_exact(f, ln).decode("utf-8"))
    if names:
        if len(names) != len(arrs):
            raise MXNetError("Invalid NDArray file format")
        return dict(zip(names, arrs))
    return arrs


_FILE_VARS = {}


def _file_var(fname, create=True):
    """Engine variable that orders asynchronous writes to (and reads of) one checkpoint file."""
    import os
    from .. import engine
    key = os.path.abspath(fname)
    v = _FILE_VARS.get(key)
    if v is None and create:
        v = _FILE_VARS[key] = engine.get().new_variable()
    return v


def save(fname, data, dtype=None, async_write=False):
    """mx.nd.save (python/mxnet/ndarray/utils.py:214).  Uses the native writer when built.

    ``async_write``: the arrays are snapshotted to host memory now (so training may overwrite them immediately) and the
    file is written by a dependency-engine worker (csrc/host/engine.cc) -- the reference pushes the same kind of host work
    through its ThreadedEngine.  ``mx.nd.load`` of that file, ``mx.nd.waitall()``, the elastic re-init drain
    (engine.stop()) and interpreter exit wait for the write."""
    from .. import _native
    nat = _native.params_io()
    if nat is not None and dtype is None and _native_ok(data):
        names, arrs = (list(data.keys()), list(data.values())) if isinstance(data, dict) else ([], list(data) if not isinstance(data, NDArray) else [data])
        if isinstance(data, dict) and not all(isinstance(k, str) for k in names):
            raise TypeError("save only accept dict str->NDArray or list of NDArray")
        staged = []
        if async_write and any(a._t.is_cuda for a in arrs):
            # device arrays are snapshotted through pinned staging blocks of the pooled storage manager: all copies are
            # queued asynchronously, ONE synchronisation covers them, and the engine worker returns the blocks to the pool
            # after the file is written (the next checkpoint finds them there)
            from .. import storage as _st
            ts = []
            for a in arrs:
                t = a._t.detach().contiguous()
                if t.is_cuda and t.numel() > 0:
                    buf = _st.alloc_pinned(t.numel() * t.element_size())
                    staged.append(buf)
                    h = buf.view(t.dtype).view(t.shape)
                    h.copy_(t, non_blocking=True)
                    ts.append(h)
                else:
                    ts.append(t.cpu().clone())
            torch.cuda.synchronize()
        else:
            ts = [a._t.detach().contiguous().cpu() for a in arrs]
            if async_write:
                ts = [t.clone() if (not a._t.is_cuda and t.data_ptr() == a._t.data_ptr()) else t for t, a in zip(ts, arrs)]
        ctxs = [(a.context.device_typeid, a.context.device_id) for a in arrs]

        def write():
            try:
                nat.save(fname, names, ts, ctxs)
            finally:
                if staged:
                    from .. import storage as _st
                    for b in staged:
                        _st.free_pinned(b)
    else:
        blob = dumps(data, dtype)

        def write():
            with open(fname, "wb") as f:
                f.write(blob)
    if async_write:
        from .. import engine
        engine.get().push(write, mutable_vars=[_file_var(fname)], name="nd.save")
        return
    v = _file_var(fname, create=False)
    if v is not None:
        from .. import engine
        engine.get().wait_for_var(v)          # an earlier asynchronous write of the same file must land first
    write()


def _native_ok(data):
    arrs = data.values() if isinstance(data, dict) else ([data] if isinstance(data, NDArray) else data)
    return all(isinstance(a, NDArray) and a._t.dim() > 0 for a in arrs)


def load(fname, ctx=None):
    """mx.nd.load."""
    v = _file_var(fname, create=False)
    if v is not None:
        from .. import engine
        engine.get().wait_for_var(v)          # asynchronous checkpoint write still in flight
    with open(fname, "rb") as f:
        return loads(f.read(), ctx)


def load_frombuffer(buf, ctx=None):
    return loads(bytes(buf), ctx)
