"""Storage manager front-end (reference: include/mxnet/storage.h ``Storage::Get()->Alloc / Free / DirectFree /
ReleaseAll``, src/storage/pooled_storage_manager.h, pinned_memory_storage.h, cpu_device_storage.h).

Two allocators coexist in this framework:

* torch tensors (activations, user NDArrays) come from torch's caching allocator -- the pooled GPU manager's role for
  framework-created arrays;
* everything the runtime allocates itself goes through the native pooled manager in ``csrc/host/storage.cc``
  (bucketed free lists per (device type, device id), "Naive" page rounding or "Round" power-of-two buckets selected by
  ``MXNET_GPU_MEM_POOL_TYPE``, the ``MXNET_GPU_MEM_POOL_RESERVE`` rule, statistics): the symmetric peer-mapped arenas of
  ``parallel/peer.py`` (an elastic re-init finds the blocks of the previous membership in the pool instead of paying
  cudaFree + cudaMalloc), pinned staging buffers of the input pipeline and of asynchronous checkpoints, kernel workspaces.
  The raw device / pinned back-ends live in the CUDA library and are registered here on first use.

``alloc(nbytes, ctx)`` returns a ``Handle`` (``.ptr``, ``.size``, ``.ctx``, ``.tensor(dtype)`` = zero-copy torch view).
"""
import ctypes
import threading

import torch

from . import _native as N
from .base import MXNetError, getenv

__all__ = ["Handle", "alloc", "free", "direct_free", "release_all", "stats", "bucket_size", "pool_type", "reserve_percent",
           "empty_cache", "memory_stats", "alloc_pinned", "free_pinned"]

CPU, GPU, PINNED = 1, 2, 3
_lock = threading.Lock()
_backends_ready = [False]


def pool_type():
    return getenv("MXNET_GPU_MEM_POOL_TYPE", "Naive")


def reserve_percent():
    return getenv("MXNET_GPU_MEM_POOL_RESERVE", 5)


def _lib():
    lib = N.host()
    if not _backends_ready[0] and torch.cuda.is_available():
        with _lock:
            if not _backends_ready[0]:
                cu = N.cuda()
                fn = lambda name: ctypes.cast(getattr(cu, name), ctypes.c_void_p)   # noqa: E731
                N.check(lib, lib.MXBStorageSetBackend(ctypes.c_int(GPU), fn("MXBRawDeviceAlloc"), fn("MXBRawDeviceFree"),
                                                      fn("MXBRawMemInfo")))
                N.check(lib, lib.MXBStorageSetBackend(ctypes.c_int(PINNED), fn("MXBRawPinnedAlloc"), fn("MXBRawPinnedFree"), None))
                _backends_ready[0] = True
    return lib


def _dev_of(ctx):
    """(device type code, device id) of a Context / torch.device / string ('cpu', 'cpu_pinned', 'gpu(1)')."""
    if ctx is None:
        return CPU, 0
    if isinstance(ctx, torch.device):
        return (GPU, ctx.index or 0) if ctx.type == "cuda" else (CPU, 0)
    if isinstance(ctx, str):
        name = ctx.split("(")[0]
        idx = int(ctx.split("(")[1].rstrip(")")) if "(" in ctx else 0
    else:
        name, idx = ctx.device_type, ctx.device_id
    return {"cpu": CPU, "gpu": GPU, "cpu_pinned": PINNED, "cpu_shared": CPU}[name], int(idx)


class Handle(object):
    """One allocation of the pooled manager (Storage::Handle: dptr, size, ctx)."""

    __slots__ = ("ptr", "size", "bucket", "dev_type", "dev_id", "_alive")

    def __init__(self, ptr, size, bucket, dev_type, dev_id):
        self.ptr, self.size, self.bucket, self.dev_type, self.dev_id, self._alive = ptr, size, bucket, dev_type, dev_id, True

    @property
    def ctx(self):
        return {CPU: "cpu(0)", GPU: "gpu(%d)" % self.dev_id, PINNED: "cpu_pinned(%d)" % self.dev_id}[self.dev_type]

    def tensor(self, dtype=torch.uint8, numel=None, offset_bytes=0):
        """Zero-copy torch view of the block (the handle must outlive it)."""
        if not self._alive:
            raise MXNetError("storage handle already freed")
        esz = torch.empty((), dtype=dtype).element_size()
        numel = (self.size - offset_bytes) // esz if numel is None else int(numel)
        if numel * esz + offset_bytes > self.bucket:
            raise MXNetError("view exceeds the allocation")
        if self.dev_type == GPU:
            from .parallel.peer import _RawBuffer
            return _RawBuffer(self.ptr, self.bucket, torch.device("cuda", self.dev_id), owner=False).tensor(dtype, numel, offset_bytes)
        buf = (ctypes.c_char * (numel * esz)).from_address(self.ptr + offset_bytes)
        t = torch.frombuffer(buf, dtype=dtype, count=numel) if numel else torch.empty(0, dtype=dtype)
        t._mxb_keepalive = self
        return t

    def __repr__(self):
        return "<storage.Handle %s %d bytes (bucket %d) at 0x%x>" % (self.ctx, self.size, self.bucket, self.ptr)


def alloc(nbytes, ctx=None):
    """``Storage::Alloc``: a block of at least ``nbytes`` on ``ctx`` (cpu / cpu_pinned / gpu), from the pool if possible."""
    lib = _lib()
    dt, di = _dev_of(ctx)
    if dt != CPU and not torch.cuda.is_available():
        raise MXNetError("storage: %s memory needs a CUDA device" % ("gpu" if dt == GPU else "pinned"))
    p, b = ctypes.c_void_p(), ctypes.c_int64()
    N.check(lib, lib.MXBStorageAlloc(ctypes.c_int(dt), ctypes.c_int(di), ctypes.c_int64(int(nbytes)), ctypes.byref(p), ctypes.byref(b)))
    return Handle(p.value, int(nbytes), b.value, dt, di)


def free(handle):
    """``Storage::Free``: back to the pool (re-used by the next allocation of the same bucket)."""
    if handle is None or not handle._alive:
        return
    lib = _lib()
    N.check(lib, lib.MXBStorageFree(ctypes.c_int(handle.dev_type), ctypes.c_int(handle.dev_id), ctypes.c_void_p(handle.ptr)))
    handle._alive = False


def direct_free(handle):
    """``Storage::DirectFree``: straight back to the device, bypassing the pool."""
    if handle is None or not handle._alive:
        return
    lib = _lib()
    N.check(lib, lib.MXBStorageDirectFree(ctypes.c_int(handle.dev_type), ctypes.c_int(handle.dev_id), ctypes.c_void_p(handle.ptr)))
    handle._alive = False


def release_all(ctx=None):
    """``Storage::ReleaseAll``: drop every cached block of this context's pool."""
    lib = _lib()
    dt, di = _dev_of(ctx)
    N.check(lib, lib.MXBStorageReleaseAll(ctypes.c_int(dt), ctypes.c_int(di)))


def stats(ctx=None):
    lib = _lib()
    dt, di = _dev_of(ctx)
    s = (ctypes.c_int64 * 6)()
    N.check(lib, lib.MXBStorageStats(ctypes.c_int(dt), ctypes.c_int(di), s))
    return {"used_bytes": s[0], "pooled_bytes": s[1], "alloc_calls": s[2], "raw_allocs": s[3], "release_all_calls": s[4],
            "live_blocks": s[5]}


def bucket_size(nbytes):
    v = ctypes.c_int64()
    lib = _lib()
    N.check(lib, lib.MXBStorageBucket(ctypes.c_int64(int(nbytes)), ctypes.byref(v)))
    return v.value


def empty_cache():
    """Everything cached anywhere: our pools and torch's caching allocator."""
    release_all("cpu")
    if torch.cuda.is_available():
        release_all("cpu_pinned")
        for i in range(torch.cuda.device_count()):
            release_all("gpu(%d)" % i)
        torch.cuda.empty_cache()


def memory_stats(device=None):
    if not torch.cuda.is_available():
        return {"allocated": 0, "reserved": 0, "total": 0, "free": 0}
    free_b, total = torch.cuda.mem_get_info(device)
    return {"allocated": torch.cuda.memory_allocated(device), "reserved": torch.cuda.memory_reserved(device),
            "total": total, "free": free_b}


# ---- pinned staging helpers (input pipeline, checkpoint snapshots) --------------------------------------------------
def alloc_pinned(nbytes):
    """uint8 tensor over a pooled pinned block (plain host memory without a GPU).  Pair with ``free_pinned``."""
    h = alloc(nbytes, "cpu_pinned" if torch.cuda.is_available() else "cpu")
    t = h.tensor(torch.uint8, numel=int(nbytes))
    t._mxb_handle = h
    return t


def free_pinned(buf):
    h = getattr(buf, "_mxb_handle", None)
    if h is not None:
        free(h)
