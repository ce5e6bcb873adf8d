"""Host -> device input pipeline stage: pinned-memory batches are copied on a side stream while the previous step
computes, and uint8 images are normalised + cast on the device in one kernel (csrc/cuda/image.cu).

Reference counterparts: ``PrefetcherIter`` / ``PrefetchingIter`` (src/io/iter_prefetcher.h, python/mxnet/io.py: a
producer thread keeps one batch ahead, in host memory) and the CPU-side normalisation of ImageRecordIter.  On B200
the copy engine + a second CUDA stream give the same one-batch-ahead overlap without a thread, and the batch crosses
the host link as uint8.

    pf = mx.io.DevicePrefetcher(source, ctx, mean=(123.68, 116.28, 103.53), std=(58.4, 57.1, 57.4), dtype="bfloat16")
    for x, y in pf:        # x: normalised NHWC bf16 NDArray on ctx, y: labels on ctx
        ...

``source`` is an iterable (or a zero-argument callable returning one) of (images, labels): torch tensors / NDArrays in
(preferably pinned) host memory; images uint8 NHWC (normalised on the device) or any float dtype (copied as is).
"""
import torch

from .. import _native as N
from ..ndarray.ndarray import NDArray
from ..ops._dtypes import code as _dt
from ..ops.functional import LAUNCHES

__all__ = ["DevicePrefetcher"]


def _t(x):
    return x._t if isinstance(x, NDArray) else x


class DevicePrefetcher(object):
    def __init__(self, source, ctx, mean=None, std=None, dtype="bfloat16", depth=2):
        self.ctx = ctx
        self.dev = ctx.torch_device
        self._src = source
        self._dtype = {"bfloat16": torch.bfloat16, "float32": torch.float32}[str(dtype)]
        self._depth = max(2, int(depth))
        self._cuda = self.dev.type == "cuda"
        self._stream = torch.cuda.Stream(self.dev) if self._cuda else None
        self._slots = [None] * self._depth            # per slot: (raw device buffers, event)
        self._mean = self._istd = None
        if mean is not None:
            self._mean = torch.tensor([float(v) for v in mean], dtype=torch.float32, device=self.dev)
            self._istd = 1.0 / torch.tensor([float(v) for v in (std if std is not None else [1.0] * len(mean))],
                                            dtype=torch.float32, device=self.dev)
        self.h2d_bytes = 0

    # ---- one batch: async copy on the side stream, normalise there too ---------------------------------------
    def _stage(self, slot, batch):
        img, lab = _t(batch[0]), _t(batch[1])
        self.h2d_bytes = img.numel() * img.element_size() + lab.numel() * lab.element_size()
        if not self._cuda:
            x = img
            if img.dtype == torch.uint8 and self._mean is not None:
                x = ((img.float() - self._mean) * self._istd).to(self._dtype)
            return (x, lab, None)
        bufs = self._slots[slot]
        if bufs is None or bufs[0].shape != img.shape or bufs[0].dtype != img.dtype or bufs[2].shape != lab.shape:
            raw = torch.empty(img.shape, dtype=img.dtype, device=self.dev)
            out = torch.empty(img.shape, dtype=self._dtype, device=self.dev) if img.dtype == torch.uint8 else raw
            lab_d = torch.empty(lab.shape, dtype=lab.dtype, device=self.dev)
            bufs = [raw, out, lab_d, torch.cuda.Event()]
            self._slots[slot] = bufs
        raw, out, lab_d, ev = bufs
        cur = torch.cuda.current_stream(self.dev)
        self._stream.wait_stream(cur)                 # the compute stream may still read this slot's previous batch
        with torch.cuda.stream(self._stream):
            raw.copy_(img, non_blocking=True)
            lab_d.copy_(lab, non_blocking=True)
            if img.dtype == torch.uint8:
                if self._mean is not None and N.have_cuda_kernels():
                    lib = N.cuda()
                    N.check(lib, lib.MXBNormalizeU8(N.ptr(raw), N.ptr(out), N.c_int(_dt(self._dtype)), N.c_int64(raw.numel()),
                                                    N.c_int(raw.shape[-1]), N.ptr(self._mean), N.ptr(self._istd),
                                                    N.c_void_p(self._stream.cuda_stream)))
                    LAUNCHES[0] += 1
                elif self._mean is not None:
                    out.copy_(((raw.float() - self._mean) * self._istd).to(self._dtype))
                else:
                    out.copy_(raw.to(self._dtype))
            ev.record(self._stream)
        return (out, lab_d, ev)

    def __iter__(self):
        it = iter(self._src() if callable(self._src) else self._src)
        slot = 0
        try:
            nxt = self._stage(slot, next(it))
        except StopIteration:
            return
        while nxt is not None:
            x, y, ev = nxt
            if ev is not None:
                torch.cuda.current_stream(self.dev).wait_event(ev)
            slot = (slot + 1) % self._depth
            try:
                nxt = self._stage(slot, next(it))     # copy of batch i+1 overlaps the compute of batch i
            except StopIteration:
                nxt = None
            yield NDArray(x, self.ctx), NDArray(y, self.ctx)
