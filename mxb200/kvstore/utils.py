"""kvstore helpers (reference: src/kvstore/kvstore_utils.{cc,cu}: UniqueImpl = radix sort + unique of the row ids).

``unique_rows(ids, num_rows)``: row ids are bounded by the parameter's row count, so on the GPU the native path
(``csrc/cuda/unique.cu``) marks them in a bitmap, scans the per-block populations and compacts -- sorted by construction,
O(n + num_rows / 32), no sort.  It is used when the bitmap is not much larger than the id list; otherwise (huge tables,
few ids), on the CPU and when ``num_rows`` is unknown, ``torch.unique`` (sort-based) is used."""
import ctypes

import torch

from .. import _native as N
from ..base import MXNetError

_scratch = {}


def unique_rows(ids, num_rows=None):
    """Sorted unique int64 row ids."""
    ids = ids.reshape(-1).to(torch.int64)
    n = ids.numel()
    if (num_rows is None or not ids.is_cuda or n == 0 or not N.have_cuda_kernels() or not hasattr(N.cuda(), "MXBUniqueRowIds")
            or num_rows > 64 * 32 * max(n, 1024)):
        return torch.unique(ids, sorted=True)
    lib = N.cuda()
    dev = ids.device
    key = (dev, int(num_rows))
    sc = _scratch.get(key)
    if sc is None:
        need = ctypes.c_int64()
        N.check(lib, lib.MXBUniqueScratchBytes(ctypes.c_int64(int(num_rows)), ctypes.byref(need)))
        sc = (torch.zeros(need.value, dtype=torch.uint8, device=dev), torch.zeros(1, dtype=torch.int64, device=dev),
              torch.zeros(1, dtype=torch.int32, device=dev))
        if len(_scratch) > 64:
            _scratch.clear()
        _scratch[key] = sc
    scratch, count, status = sc
    ids = ids.contiguous()
    out = torch.empty(min(n, int(num_rows)), dtype=torch.int64, device=dev)
    N.check(lib, lib.MXBUniqueRowIds(N.ptr(ids), ctypes.c_int64(n), ctypes.c_int64(int(num_rows)), N.ptr(out), N.ptr(scratch), N.ptr(count),
                                     N.ptr(status), N.stream_of(ids)))
    from ..ops.functional import LAUNCHES
    LAUNCHES[0] += 4
    cnt, bad = int(count.item()), int(status.item())          # one synchronisation: the caller needs the length anyway
    if bad:
        status.zero_()
        raise MXNetError("row id out of range [0, %d)" % int(num_rows))
    return out[:cnt]
