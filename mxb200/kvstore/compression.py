"""2-bit gradient compression with error-feedback residual (reference: src/kvstore/gradient_compression.{h,cc},
-inl.h:40-127).

Wire format (kept bit-exact so the reference's NumPy oracle, tests/nightly/test_kvstore.py:33-77, applies):
16 gradient values -> one 32-bit word; value i of a block lives in byte (i>>2), bit pair (i&3) counted from
the MSB: ``11`` = +threshold, ``10`` = -threshold, ``00`` = 0.  residual += grad; emit +t when residual >= t
(residual -= t), emit -t when residual <= -t (residual += t).  Compressed size = ceil(n/16) words.
"""
import torch

from ..base import MXNetError

_POS = (0xC0, 0x30, 0x0C, 0x03)
_NEG = (0x80, 0x20, 0x08, 0x02)


class GradientCompression(object):
    def __init__(self, params):
        params = dict(params or {})
        self.type = params.get("type", "2bit")
        if self.type not in ("2bit", "none"):
            raise MXNetError("Unknown type for gradient compression " + str(self.type))
        self.threshold = float(params.get("threshold", 0.5))
        if self.threshold <= 0:
            raise MXNetError("threshold must be greater than 0 for two bit compression")

    def encode_params(self):
        """'type,threshold' string sent to servers (gradient_compression.cc:60-75)."""
        return "%d,%s" % (1 if self.type == "2bit" else 0, repr(self.threshold))

    @staticmethod
    def compressed_size(n):
        return (n + 15) // 16

    @staticmethod
    def _native(t):
        if not t.is_cuda:
            return None
        from .. import _native as N
        return N.cuda() if N.have_cuda_kernels() else None

    def quantize(self, grad, residual):
        """grad (any float dtype), residual (fp32, updated in place) -> int32 words [ceil(n/16)]."""
        t = self.threshold
        n = grad.numel()
        lib = self._native(grad)
        if lib is not None and grad.dtype in (torch.float32, torch.bfloat16) and grad.is_contiguous() \
                and residual.is_contiguous() and residual.dtype is torch.float32:
            # one pass over grad + residual on the device (csrc/cuda/compress.cu)
            from .. import _native as N
            from ..ops._dtypes import code as _dt
            from ..ops.functional import LAUNCHES
            words = torch.empty(self.compressed_size(n), dtype=torch.int32, device=grad.device)
            N.check(lib, lib.MXBQuantize2Bit(N.ptr(grad), N.c_int(_dt(grad.dtype)), N.ptr(residual), N.ptr(words),
                                             N.c_int64(n), N.c_float(t), N.stream_of(grad)))
            LAUNCHES[0] += 1
            return words
        with torch.no_grad():
            residual.add_(grad.reshape(residual.shape).float())
            r = residual.reshape(-1)
            pos = r >= t
            neg = r <= -t
            r.sub_(pos.float() * t)
            r.add_(neg.float() * t)
            pad = (-n) % 16
            code = (pos.to(torch.int64) * 3 + neg.to(torch.int64) * 2)   # 3 = '11', 2 = '10'
            if pad:
                code = torch.cat([code, torch.zeros(pad, dtype=torch.int64, device=code.device)])
            code = code.reshape(-1, 4, 4)          # [word, byte, col]
            shifts = torch.tensor([6, 4, 2, 0], device=code.device)
            bytes_ = (code << shifts).sum(-1)      # [word, 4] each 0..255
            bshift = torch.tensor([0, 8, 16, 24], device=code.device)   # little-endian: byte 0 is the LSB
            words = (bytes_ << bshift).sum(-1)
            words = torch.where(words >= 2 ** 31, words - 2 ** 32, words)
            return words.to(torch.int32)

    def dequantize(self, words, n, shape, device=None, out=None):
        """words -> fp32 tensor of ``shape``; with ``out`` (fp32, contiguous) the decoded values are ADDED into it
        (how a server merges the compressed pushes of several workers)."""
        t = self.threshold
        lib = self._native(words)
        if lib is not None and words.dtype is torch.int32 and words.is_contiguous():
            from .. import _native as N
            from ..ops.functional import LAUNCHES
            acc = out is not None
            res = out if acc else torch.empty(n, dtype=torch.float32, device=words.device)
            N.check(lib, lib.MXBDequantize2Bit(N.ptr(words), N.ptr(res), N.c_int64(n), N.c_float(t),
                                               N.c_int(1 if acc else 0), N.stream_of(words)))
            LAUNCHES[0] += 1
            return res.reshape(shape)
        if out is not None:
            out.add_(self.dequantize(words, n, out.shape, device))
            return out
        with torch.no_grad():
            w = words.to(torch.int64) & 0xFFFFFFFF
            bshift = torch.tensor([0, 8, 16, 24], device=w.device)
            bytes_ = (w.unsqueeze(-1) >> bshift) & 0xFF
            shifts = torch.tensor([6, 4, 2, 0], device=w.device)
            code = (bytes_.unsqueeze(-1) >> shifts) & 0x3
            code = code.reshape(-1)[:n]
            out = torch.zeros(n, dtype=torch.float32, device=w.device)
            out[code == 3] = t
            out[code == 2] = -t
            return out.reshape(shape)
