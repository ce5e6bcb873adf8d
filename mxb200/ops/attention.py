"""Fused multi-head self-attention (csrc/cuda/attention.cu): scores, key-padding mask, softmax, dropout and the value
product of one (batch, head) in one tcgen05 kernel; backward recomputes the probabilities from the saved log-sum-exp.

The reference has no attention operator -- a BERT on it composes ``batch_dot`` + ``softmax`` + ``Dropout`` + ``batch_dot``
(src/operator/tensor/dot-inl.cuh, src/operator/nn/softmax.cc, src/operator/nn/dropout-inl.h).

    out = self_attention(q, k, v, num_heads, key_mask=None, p_drop=0.0, training=True)

q, k, v: [B, S, H] (the projection outputs, head h in columns [64h, 64h + 64)); out: [B, S, H].  The kernel covers
S <= 128 and head dim 64 in bf16; anything else takes the eager composition (and on a GPU host without the native library
this raises instead of silently falling back).
"""
import ctypes
import math

import torch

from .. import _native as N
from ..base import getenv
from .functional import LAUNCHES

c_void_p, c_int, c_int64, c_float, c_uint64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_uint64

_STEP = {}      # device -> int64 device tensor: Philox stream offset, advanced once per training forward (graph safe)


def _p(t):
    return c_void_p(t.data_ptr()) if t is not None else c_void_p(0)


def rng_step_counter(device):
    """Device-resident counter added to every attention call's Philox offset.  ``advance_rng_step`` bumps it with a device
    op, so a captured CUDA graph draws fresh dropout masks on every replay."""
    t = _STEP.get(device)
    if t is None:
        t = _STEP[device] = torch.zeros(1, dtype=torch.int64, device=device)
    return t


def advance_rng_step(device):
    rng_step_counter(device).add_(1 << 20)


def attention_ok(q, num_heads):
    if not (q.is_cuda and q.dtype is torch.bfloat16 and q.dim() == 3 and N.have_cuda_kernels()):
        return False
    if getenv("MXB200_ATTENTION", "tcgen05") != "tcgen05":
        return False
    B, S, H = q.shape
    return S <= 128 and H == num_heads * 64 and q.stride(2) == 1 and q.stride(1) % 8 == 0 and q.stride(0) == S * q.stride(1)


class _AttnFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, key_mask, nh, scale, p_drop, call_id):
        B, S, H = q.shape
        ld = q.stride(1)
        assert k.stride() == q.stride() and v.stride() == q.stride(), "q, k, v must share one layout"
        out = torch.empty((B, S, H), dtype=torch.bfloat16, device=q.device)
        lse = torch.empty((B, nh, S), dtype=torch.float32, device=q.device)
        rng = torch.empty(2, dtype=torch.int64, device=q.device)
        seed = int(torch.initial_seed()) & 0x7FFFFFFFFFFFFFFF
        lib = N.cuda()
        N.check(lib, lib.MXBAttnFwd(_p(q), _p(k), _p(v), _p(out), _p(lse), c_int(B), c_int(S), c_int(nh), c_int64(ld),
                                    c_float(scale), _p(key_mask), c_float(p_drop), c_uint64(seed), c_uint64(int(call_id)),
                                    _p(rng_step_counter(q.device)) if p_drop > 0 else c_void_p(0), _p(rng),
                                    N.stream_of(q)))
        LAUNCHES[0] += 1
        ctx.save_for_backward(q, k, v, out, lse, rng, key_mask)
        ctx.cfg = (nh, scale, p_drop)
        return out

    @staticmethod
    def backward(ctx, dout):
        q, k, v, out, lse, rng, key_mask = ctx.saved_tensors
        nh, scale, p_drop = ctx.cfg
        B, S, H = q.shape
        if not (dout.is_contiguous()):
            dout = dout.contiguous()
        # one [3, B, S, H] buffer: dq / dk / dv are row-contiguous [B*S, H] matrices the projection wgrad GEMMs read directly
        g = torch.empty((3, B, S, H), dtype=torch.bfloat16, device=q.device)
        lib = N.cuda()
        if q.stride(1) != H:
            q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
        N.check(lib, lib.MXBAttnBwd(_p(q), _p(k), _p(v), _p(out), _p(dout), _p(lse), _p(g[0]), _p(g[1]), _p(g[2]), c_int(B),
                                    c_int(S), c_int(nh), c_int64(H), c_float(scale), _p(key_mask), c_float(p_drop), _p(rng),
                                    N.stream_of(q)))
        LAUNCHES[0] += 1
        return g[0], g[1], g[2], None, None, None, None, None


def self_attention(q, k, v, num_heads, key_mask=None, p_drop=0.0, training=True, scale=None, call_id=0):
    """Multi-head scaled-dot-product self-attention.  ``key_mask``: [B, S] bool / uint8, True = attend."""
    B, S, H = q.shape
    dh = H // num_heads
    scale = float(scale) if scale is not None else 1.0 / math.sqrt(dh)
    p = float(p_drop) if training else 0.0
    if attention_ok(q, num_heads) and k.stride() == q.stride() and v.stride() == q.stride() and k.dtype is q.dtype:
        m = None
        if key_mask is not None:
            m = key_mask.to(torch.uint8).contiguous()
        return _AttnFn.apply(q, k, v, m, num_heads, scale, p, call_id)
    if q.is_cuda and not N.have_cuda_kernels():
        raise RuntimeError("mxb200: CUDA kernels library missing on a GPU host")
    # eager composition (CPU, long sequences, other head sizes)
    qh = q.reshape(B, S, num_heads, dh).transpose(1, 2)
    kh = k.reshape(B, S, num_heads, dh).transpose(1, 2)
    vh = v.reshape(B, S, num_heads, dh).transpose(1, 2)
    s = torch.matmul(qh.float(), kh.float().transpose(-1, -2)) * scale
    if key_mask is not None:
        s = s.masked_fill(~key_mask.to(torch.bool)[:, None, None, :], float("-inf"))
    a = torch.softmax(s, dim=-1)
    if p > 0:
        a = torch.nn.functional.dropout(a, p)
    return torch.matmul(a, vh.float()).to(q.dtype).transpose(1, 2).reshape(B, S, H)
