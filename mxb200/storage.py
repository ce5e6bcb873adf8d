"""Storage manager front-end (reference: include/mxnet/storage.h, src/storage/pooled_storage_manager.h).

Device memory pooling is torch's caching allocator (the pooled GPU manager's role); this module exposes the
reference's knobs and statistics, the pinned-host pool used by IO/checkpoint staging, and the symmetric
(peer-mapped) allocations of mxb200.parallel.peer which are cudaMalloc'ed directly so they can be exported
over CUDA IPC."""
import torch

from .base import getenv

__all__ = ["pool_type", "reserve_percent", "empty_cache", "memory_stats", "alloc_pinned"]

_pinned_free = {}


def pool_type():
    return getenv("MXNET_GPU_MEM_POOL_TYPE", "Naive")


def reserve_percent():
    return getenv("MXNET_GPU_MEM_POOL_RESERVE", 5)


def empty_cache():
    """Storage::ReleaseAll analogue."""
    _pinned_free.clear()
    if torch.cuda.is_available():
        torch.cuda.empty_cache()


def memory_stats(device=None):
    if not torch.cuda.is_available():
        return {"allocated": 0, "reserved": 0, "total": 0, "free": 0}
    free, total = torch.cuda.mem_get_info(device)
    return {"allocated": torch.cuda.memory_allocated(device), "reserved": torch.cuda.memory_reserved(device),
            "total": total, "free": free}


def alloc_pinned(nbytes):
    """Rounded (power-of-two) pinned host buffers with reuse (GPUPooledRoundedStorageManager bucketing,
    applied to cudaMallocHost memory)."""
    size = 1 << max(12, (int(nbytes) - 1).bit_length())
    lst = _pinned_free.get(size)
    if lst:
        return lst.pop()
    return torch.empty(size, dtype=torch.uint8, pin_memory=torch.cuda.is_available())


def free_pinned(buf):
    _pinned_free.setdefault(buf.numel(), []).append(buf)
