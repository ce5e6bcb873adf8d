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

    def quantize(self, grad, residual):
        """grad (any float dtype), residual (fp32, updated in place) -> int32 words [ceil(n/16)]."""
        t = self.threshold
        n = grad.numel()
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

    def dequantize(self, words, n, shape, device=None):
        t = self.threshold
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
