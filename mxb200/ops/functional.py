"""Torch-level functional ops backed by the hand-written sm_100a kernels.

Every function takes/returns ``torch.Tensor`` and is differentiable through a
``torch.autograd.Function`` whose forward AND backward call our kernels:

  linear / conv1x1_nhwc -> csrc/cuda/gemm_tcgen05.cu  (TMA + tcgen05.mma + TMEM)
  bn_act                -> csrc/cuda/bn_act.cu        (BatchNorm + residual + ReLU, NHWC)
  softmax_cross_entropy -> csrc/cuda/softmax_ce.cu    (loss + gradient in one launch)

CPU tensors (and shapes the kernels do not cover) take an equivalent eager
torch path so the whole framework runs on a GPU-less host; on a CUDA device
a missing native library is an error, never a silent fallback.
"""
import ctypes

import torch
import torch.nn.functional as F

from .. import _native as N
from ..base import getenv
from ._dtypes import code

c_void_p, c_int, c_int64, c_float = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float

LAUNCHES = [0]


def _p(t):
    return c_void_p(t.data_ptr()) if t is not None else c_void_p(0)


def _use_native(*ts):
    return all(t is None or t.is_cuda for t in ts) and N.have_cuda_kernels()


# --------------------------------------------------------------------------- GEMM
GEMM_NT, GEMM_NN, GEMM_TN = 0, 1, 2
STATS_OFF = 8     # BN workspace layout: [counter | pad] [sums | sumsq] ... (csrc/cuda/bn_act.cu)


def gemm_raw(mode, A, B, D, M, Nn, K, lda, ldb, ldd, bias=None, relu=False, col_sum=None, col_sumsq=None, addend=None,
             ldc=0, splits=1, block_n=0, max_ctas=0, dbg=None):
    lib = N.cuda()
    dbg_arr = (c_int * 4)(*dbg) if dbg is not None else None
    rc = lib.MXBGemmBf16(c_int(mode), _p(A), _p(B), _p(D), c_int(M), c_int(Nn), c_int(K), c_int64(lda), c_int64(ldb),
                         c_int64(ldd), _p(bias), c_int(1 if relu else 0), _p(col_sum), _p(col_sumsq), _p(addend),
                         c_int64(ldc), c_int(splits), c_int(block_n), c_int(max_ctas), dbg_arr, N.stream_of(A))
    N.check(lib, rc)
    LAUNCHES[0] += 1


def gemm_bias_act(x, w, bias=None, relu=False, addend=None, block_n=0):
    """Inference epilogue: y = relu(x.W^T + bias + addend) in one kernel (no autograd)."""
    M, K = x.shape
    Nn = w.shape[0]
    if not (_gemm_ok(x, w) and (addend is None or (addend.is_contiguous() and addend.dtype is torch.bfloat16))):
        y = F.linear(x, w, bias.to(x.dtype) if bias is not None else None)
        if addend is not None:
            y = y + addend
        return torch.relu(y) if relu else y
    if not x.is_contiguous():
        x = x.contiguous()
    y = torch.empty((M, Nn), dtype=torch.bfloat16, device=x.device)
    gemm_raw(GEMM_NT, x, w, y, M, Nn, K, x.stride(0), w.stride(0), Nn, bias, relu, addend=addend,
             ldc=addend.stride(0) if addend is not None else 0, block_n=block_n)
    return y


def gemm_nt(x, w, bias=None, relu=False, col_stats=None, block_n=0):
    """y[M,N] = x[M,K] @ w[N,K]^T (+bias)(relu); bf16 in/out, fp32 accumulate in TMEM.  ``col_stats`` =
    (sum, sumsq) fp32 [N] accumulators that receive the per-column statistics of y (fused BatchNorm stats)."""
    M, K = x.shape
    Nn = w.shape[0]
    y = torch.empty((M, Nn), dtype=torch.bfloat16, device=x.device)
    cs, cs2 = col_stats if col_stats is not None else (None, None)
    gemm_raw(GEMM_NT, x, w, y, M, Nn, K, x.stride(0), w.stride(0), Nn, bias, relu, cs, cs2, block_n=block_n)
    return y


def gemm_nn(dy, w, addend=None, block_n=0):
    """dx[M,K] = dy[M,N] @ w[N,K] (+ addend[M,K]: the residual-branch gradient, added in the epilogue)."""
    M, Nn = dy.shape
    K = w.shape[1]
    dx = torch.empty((M, K), dtype=torch.bfloat16, device=dy.device)
    gemm_raw(GEMM_NN, dy, w, dx, M, K, Nn, dy.stride(0), w.stride(0), K, addend=addend,
             ldc=addend.stride(0) if addend is not None else 0, block_n=block_n)
    return dx


def _pick_splits(M_out, N_out, K_red, block_n):
    tiles = ((M_out + 127) // 128) * ((N_out + block_n - 1) // block_n)
    kb = (K_red + 63) // 64
    want = max(1, (148 * 2) // max(1, tiles))
    return max(1, min(want, kb // 4 if kb >= 8 else 1))


def gemm_tn(dy, x, out_dtype=torch.bfloat16, block_n=0, out=None):
    """dw[N,K] = dy[M,N]^T @ x[M,K]; reduction over the (huge) M dim is split across CTAs and combined with
    fp32 TMA reduce-add.  ``out``: fp32 [N,K] buffer that is ACCUMULATED into (e.g. a gradient-arena view)."""
    M, Nn = dy.shape
    K = x.shape[1]
    bn = block_n or (64 if K <= 64 else (128 if K < 256 else 256))   # measured: 256-wide tiles win for K >= 256 (profiles/)
    acc = out if out is not None else torch.zeros((Nn, K), dtype=torch.float32, device=dy.device)
    gemm_raw(GEMM_TN, dy, x, acc, Nn, K, M, dy.stride(0), x.stride(0), acc.stride(0), splits=_pick_splits(Nn, K, M, bn),
             block_n=bn)
    if out is not None:
        return None
    return acc if out_dtype is torch.float32 else acc.to(out_dtype)


def _gemm_ok(x, w):
    if not (x.is_cuda and x.dtype is torch.bfloat16 and w.dtype is torch.bfloat16):
        return False
    if getenv("MXB200_GEMM", "tcgen05") != "tcgen05":
        return False
    K, Nn = x.shape[-1], w.shape[0]
    return K % 8 == 0 and Nn % 8 == 0 and x.stride(-1) == 1 and w.stride(-1) == 1 and N.have_cuda_kernels()


def sink_touch(*sinks):
    """'write' gradient semantics for gradient-arena sinks (reference: grad_req='write', every backward pass overwrites).
    The arena is accumulated into by the backward kernels and cleared by the fused optimizer step; if a backward pass is
    NOT followed by that trainer's step (GAN-style loops, skipped steps, repeated backward calls) the next backward that
    touches the arena clears the stale gradients first -- one memset per arena and backward pass, only when needed."""
    from .. import autograd as _ag
    ep = _ag._bw_epoch[0]
    for sk in sinks:
        ar = getattr(sk, "_mxb_arena", None) if sk is not None else None
        if ar is None or ar.bw_epoch == ep:
            continue
        if ar.dirty and ar.write_semantics:
            ar.grads.zero_()
        ar.bw_epoch = ep
        ar.dirty = True


def _sink_ok(sink, shape):
    return sink is not None and sink.dtype is torch.float32 and tuple(sink.shape) == tuple(shape) and sink.is_contiguous()


def colsum(x2, out=None):
    """fp32 column sums of a contiguous [rows, C] tensor; ``out`` (fp32 [C]) is accumulated into (gradient-arena view)."""
    C = x2.shape[-1]
    if _use_native(x2) and x2.dim() == 2 and x2.is_contiguous() and C % 8 == 0 and x2.dtype in (torch.bfloat16, torch.float32) \
            and (out is None or (out.dtype is torch.float32 and out.is_contiguous())):
        res = out if out is not None else torch.zeros(C, dtype=torch.float32, device=x2.device)
        lib = N.cuda()
        N.check(lib, lib.MXBColSumAccumulate(_p(x2), c_int(code(x2.dtype)), _p(res), c_int64(x2.shape[0]), c_int(C),
                                             N.stream_of(x2)))
        LAUNCHES[0] += 1
        return res
    s_ = x2.sum(0, dtype=torch.float32)
    if out is not None:
        out.add_(s_)
        return out
    return s_


class _LinearFn(torch.autograd.Function):
    """y = x.W^T (+b)(relu).  Optional extras (all used by the fused ResNet):
      * ``stats``   : fp32 workspace receiving column sum / sum-of-squares of y (BN statistics for free);
      * ``sink_w`` / ``sink_b``: fp32 gradient-arena views; when given the weight / bias gradients are accumulated
        straight into them by the kernels and autograd receives None for those inputs;
      * ``passthrough``: also return x itself so that a residual branch can hang off the *output* of this op; the
        gradient arriving on that second output is added inside the dgrad GEMM epilogue (no separate add kernel)."""

    @staticmethod
    def forward(ctx, x, w, b, relu, stats, sink_w, sink_b, passthrough):
        x2 = x.reshape(-1, x.shape[-1])
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        bias32 = b.float() if b is not None else None
        Nn = w.shape[0]
        cs = (stats[STATS_OFF:STATS_OFF + Nn], stats[STATS_OFF + Nn:STATS_OFF + 2 * Nn]) if stats is not None else None
        y = gemm_nt(x2, w, bias32, relu, col_stats=cs)
        ctx.save_for_backward(x2, w, y if relu else None)
        ctx.has_bias = b is not None
        ctx.relu = relu
        ctx.xshape = x.shape
        ctx.bdtype = b.dtype if b is not None else None
        ctx.sink_w = sink_w if _sink_ok(sink_w, w.shape) else None
        ctx.sink_b = sink_b if (b is not None and _sink_ok(sink_b, b.shape)) else None
        out = y.reshape(x.shape[:-1] + (Nn,))
        if passthrough:
            return out, x
        return out

    @staticmethod
    def backward(ctx, gy, gskip=None):
        x2, w, y = ctx.saved_tensors
        sink_touch(ctx.sink_w, ctx.sink_b)
        g2 = gy.reshape(-1, gy.shape[-1])
        if ctx.relu:
            g2 = g2 * (y > 0)
        if not g2.is_contiguous():
            g2 = g2.contiguous()
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            add = None
            if gskip is not None:
                add = gskip.reshape(-1, gskip.shape[-1])
                if not add.is_contiguous():
                    add = add.contiguous()
            gx = gemm_nn(g2, w, addend=add).reshape(ctx.xshape)
        if ctx.needs_input_grad[1] or ctx.sink_w is not None:
            if ctx.sink_w is not None:
                gemm_tn(g2, x2, out=ctx.sink_w)
            else:
                gw = gemm_tn(g2, x2, out_dtype=w.dtype)
        if ctx.has_bias and (ctx.needs_input_grad[2] or ctx.sink_b is not None):
            if ctx.sink_b is not None:
                colsum(g2, out=ctx.sink_b)               # one kernel, straight into the fp32 gradient arena
            else:
                gb = colsum(g2).to(ctx.bdtype)
        return gx, gw, gb, None, None, None, None, None


def linear(x, w, b=None, relu=False, stats=None, sink_w=None, sink_b=None, passthrough=False):
    """FullyConnected: Y = X.W^T + b (reference: src/operator/nn/fully_connected-inl.h:72-180)."""
    if _gemm_ok(x, w):
        return _LinearFn.apply(x, w, b, relu, stats, sink_w, sink_b, passthrough)
    if x.is_cuda and not N.have_cuda_kernels():
        raise RuntimeError("mxb200: CUDA kernels library missing on a GPU host")
    y = F.linear(x, w.to(x.dtype) if w.dtype != x.dtype else w, b.to(x.dtype) if b is not None and b.dtype != x.dtype else b)
    y = torch.relu(y) if relu else y
    return (y, x) if passthrough else y


class _SubsampleFn(torch.autograd.Function):
    """x[:, ::s, ::s, :] on a contiguous NHWC tensor; the backward writes the whole input gradient in one pass
    (csrc/cuda/pool.cu) instead of zero-fill + strided scatter."""

    @staticmethod
    def forward(ctx, x, s):
        n, h, w, c = x.shape
        ho, wo = (h + s - 1) // s, (w + s - 1) // s
        y = torch.empty((n, ho, wo, c), dtype=x.dtype, device=x.device)
        lib = N.cuda()
        N.check(lib, lib.MXBSubsampleNhwc(_p(x), _p(y), c_int(x.element_size()), c_int(n), c_int(h), c_int(w), c_int(c), c_int(s),
                                          c_int(0), N.stream_of(x)))
        LAUNCHES[0] += 1
        ctx.meta = (x.shape, s)
        return y

    @staticmethod
    def backward(ctx, dy):
        (n, h, w, c), s = ctx.meta
        dy = dy.contiguous()
        dx = torch.empty((n, h, w, c), dtype=dy.dtype, device=dy.device)
        lib = N.cuda()
        N.check(lib, lib.MXBSubsampleNhwc(_p(dy), _p(dx), c_int(dy.element_size()), c_int(n), c_int(h), c_int(w), c_int(c),
                                          c_int(s), c_int(1), N.stream_of(dy)))
        LAUNCHES[0] += 1
        return dx, None


def subsample_nhwc(x, s):
    if s == 1:
        return x
    if _use_native(x) and x.dim() == 4 and x.is_contiguous() and (x.shape[-1] * x.element_size()) % 16 == 0:
        return _SubsampleFn.apply(x, int(s))
    return x[:, ::s, ::s, :].contiguous()


def conv1x1_nhwc(x, w, stride=1, stats=None, sink_w=None, passthrough=False):
    """1x1 convolution on a channels-last (N,H,W,C) tensor == GEMM over pixels.  w: (Cout, Cin)."""
    if stride != 1:
        x = subsample_nhwc(x, stride)
    n, h, ww, c = x.shape
    r = linear(x.reshape(n * h * ww, c), w, None, False, stats, sink_w, None, passthrough)
    if passthrough:
        y, xp = r
        return y.reshape(n, h, ww, w.shape[0]), xp.reshape(n, h, ww, c)
    return r.reshape(n, h, ww, w.shape[0])


# --------------------------------------------------------------------------- k x k convolution: implicit GEMM on tcgen05
def _krsc(w):
    """[O, kh, kw, I] contiguous view of a (O, I, kh, kw) weight stored channels-last (KRSC), or None."""
    v = w.permute(0, 2, 3, 1)
    return v if v.is_contiguous() else None


def conv_nhwc_ok(x, w):
    """Can csrc/cuda/gemm_tcgen05.cu run this convolution as an implicit GEMM (im2col-mode TMA)?"""
    if not (x.is_cuda and x.dim() == 4 and w.dim() == 4 and x.dtype is torch.bfloat16 and w.dtype is torch.bfloat16):
        return False
    if getenv("MXB200_CONV", "tcgen05") != "tcgen05" or not N.have_cuda_kernels():
        return False
    o, c, kh, kw = w.shape
    return (c % 64 == 0 and o % 64 == 0 and kh * kw <= 49 and x.shape[-1] == c and _krsc(w) is not None
            and x.numel() < 2 ** 31)


def conv_fwd_raw(x, wk, stride, pad, bias=None, relu=False, col_stats=None, block_n=0, addend=None):
    n, h, ww, c = x.shape
    o, kh, kw, _ = wk.shape
    P, Q = (h + 2 * pad - kh) // stride + 1, (ww + 2 * pad - kw) // stride + 1
    y = torch.empty((n, P, Q, o), dtype=torch.bfloat16, device=x.device)
    cs, cs2 = col_stats if col_stats is not None else (None, None)
    lib = N.cuda()
    N.check(lib, lib.MXBConvFwdBf16(_p(x), _p(wk), _p(y), c_int(n), c_int(h), c_int(ww), c_int(c), c_int(o), c_int(kh), c_int(kw),
                                    c_int(stride), c_int(pad), _p(bias), c_int(1 if relu else 0), _p(cs), _p(cs2),
                                    _p(addend), c_int(block_n), N.stream_of(x)))
    LAUNCHES[0] += 1
    return y


def conv_dgrad_raw(dy, wk, xshape, stride, pad, addend=None):
    n, h, ww, c = xshape
    o, kh, kw, _ = wk.shape
    dx = torch.empty((n, h, ww, c), dtype=torch.bfloat16, device=dy.device)
    lib = N.cuda()
    N.check(lib, lib.MXBConvDgradBf16(_p(dy), _p(wk), _p(dx), c_int(n), c_int(h), c_int(ww), c_int(c), c_int(o), c_int(kh),
                                      c_int(kw), c_int(stride), c_int(pad), _p(addend), N.stream_of(dy)))
    LAUNCHES[0] += 1 if stride == 1 else stride * stride
    return dx


def conv_wgrad_raw(dy, x, kshape, stride, pad, out=None, splits=0, block_n=0):
    """dW (fp32, [O, kh, kw, I]) += dy^T . im2col(x).  ``out``: contiguous fp32 buffer that is accumulated into."""
    n, h, ww, c = x.shape
    o, kh, kw, _ = kshape
    acc = out if out is not None else torch.zeros((o, kh, kw, c), dtype=torch.float32, device=x.device)
    lib = N.cuda()
    N.check(lib, lib.MXBConvWgradBf16(_p(dy), _p(x), _p(acc), c_int(n), c_int(h), c_int(ww), c_int(c), c_int(o), c_int(kh),
                                      c_int(kw), c_int(stride), c_int(pad), c_int(splits), c_int(block_n), N.stream_of(x)))
    LAUNCHES[0] += 1
    return acc


class _ConvNhwcFn(torch.autograd.Function):
    """y = conv(x, w) on channels-last tensors, all three passes on the implicit-GEMM tcgen05 kernel.  ``stats``: BN
    workspace receiving the per-channel sum / sum-of-squares of y (GEMM epilogue); ``sink_w``: fp32 gradient-arena view of
    the weight (same KRSC strides) that receives dW directly."""

    @staticmethod
    def forward(ctx, x, w, stride, pad, stats, sink_w):
        if not x.is_contiguous():
            x = x.contiguous()
        wk = _krsc(w)
        o = wk.shape[0]
        cs = (stats[STATS_OFF:STATS_OFF + o], stats[STATS_OFF + o:STATS_OFF + 2 * o]) if stats is not None else None
        y = conv_fwd_raw(x, wk, stride, pad, col_stats=cs)
        ctx.save_for_backward(x, w)
        ctx.cfg = (stride, pad)
        sk = _krsc(sink_w) if (sink_w is not None and sink_w.dtype is torch.float32 and sink_w.shape == w.shape) else None
        ctx.sink = sk
        ctx.sink_owner = sink_w if sk is not None else None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        stride, pad = ctx.cfg
        sink_touch(ctx.sink_owner)
        wk = _krsc(w)
        if not dy.is_contiguous():
            dy = dy.contiguous()
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = conv_dgrad_raw(dy, wk, x.shape, stride, pad)
        if ctx.sink is not None:
            conv_wgrad_raw(dy, x, wk.shape, stride, pad, out=ctx.sink)
        elif ctx.needs_input_grad[1]:
            dw = conv_wgrad_raw(dy, x, wk.shape, stride, pad).permute(0, 3, 1, 2).to(w.dtype)
        return dx, dw, None, None, None, None


def conv_nhwc(x, w, stride=1, pad=0, stats=None, sink_w=None):
    """k x k convolution of an NHWC bf16 tensor with a (O, I, kh, kw) weight stored KRSC.  Reference operator:
    Convolution (src/operator/nn/convolution.cu:40-128; cuDNN or im2col+GEMM there)."""
    if conv_nhwc_ok(x, w):
        return _ConvNhwcFn.apply(x, w, int(stride), int(pad), stats, sink_w)
    if x.is_cuda and not N.have_cuda_kernels():
        raise RuntimeError("mxb200: CUDA kernels library missing on a GPU host")
    return F.conv2d(x.permute(0, 3, 1, 2), w, None, stride, pad).permute(0, 2, 3, 1)


class _StemConvFn(torch.autograd.Function):
    """Small-channel k x k convolution (the ResNet stem) as im2col + tcgen05 GEMM (csrc/cuda/im2col.cu).
    The input image gets no gradient (it is data)."""

    @staticmethod
    def forward(ctx, x, w, stride, pad, stats, sink_w, bias=None, relu=False):
        n, h, ww, c = x.shape
        o, _, kh, kw = w.shape
        oh, ow = (h + 2 * pad - kh) // stride + 1, (ww + 2 * pad - kw) // stride + 1
        kwp = kw + (kw & 1)                 # patch layout (kh, kw padded to even, channel padded to 4)
        kp = kh * kwp * 4
        m = n * oh * ow
        x = x.contiguous()
        a = torch.empty((m, kp), dtype=torch.bfloat16, device=x.device)
        lib = N.cuda()
        N.check(lib, lib.MXBIm2colNHWC(_p(x), _p(a), c_int(n), c_int(h), c_int(ww), c_int(c), c_int(kh), c_int(kw),
                                       c_int(stride), c_int(pad), c_int(oh), c_int(ow), c_int(kwp), N.stream_of(x)))
        LAUNCHES[0] += 1
        wp = torch.zeros((o, kh, kwp, 4), dtype=torch.bfloat16, device=x.device)
        wp[:, :, :kw, :c] = w.permute(0, 2, 3, 1)              # (O, kh, kw, C) order == the patch order
        wp = wp.view(o, kp)
        cs = (stats[STATS_OFF:STATS_OFF + o], stats[STATS_OFF + o:STATS_OFF + 2 * o]) if stats is not None else None
        y = gemm_nt(a, wp, bias=bias, relu=relu, col_stats=cs)
        ctx.save_for_backward(a)
        ctx.meta = (w.shape, w.dtype, kwp, kp, sink_w)
        return y.view(n, oh, ow, o)

    @staticmethod
    def backward(ctx, dy):
        (a,) = ctx.saved_tensors
        wshape, wdtype, kwp, kp, sink_w = ctx.meta
        sink_touch(sink_w)
        o, c, kh, kw = wshape
        dy2 = dy.reshape(-1, o)
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        dwp = gemm_tn(dy2, a, out_dtype=torch.float32)                      # [O, Kp] fp32
        dw = dwp.view(o, kh, kwp, 4)[:, :, :kw, :c].permute(0, 3, 1, 2)     # view in (O, C, kh, kw) index order
        if sink_w is not None:
            sink_w.add_(dw)
            return None, None, None, None, None, None, None, None
        return None, dw.to(wdtype), None, None, None, None, None, None


def stem_conv_nhwc(x, w, stride, pad, stats=None, sink_w=None, bias=None, relu=False):
    """k x k convolution with very few input channels on an NHWC bf16 image.  ``stats``: BN workspace that receives
    the per-channel sum / sum-of-squares of the output (GEMM epilogue); ``bias`` / ``relu``: inference epilogue (not
    differentiated)."""
    return _StemConvFn.apply(x, w, stride, pad, stats, sink_w, bias, relu)


def stem_conv_ok(x, w):
    o, c, kh, kw = w.shape
    return (x.is_cuda and x.dtype is torch.bfloat16 and w.dtype is torch.bfloat16 and not x.requires_grad
            and c <= 4 and o % 8 == 0 and kh * (kw + 1) * 4 <= 2048 and N.have_cuda_kernels()
            and getenv("MXB200_STEM", "im2col") == "im2col")


# --------------------------------------------------------------------------- BatchNorm(+add)(+ReLU)
class _BnWorkspace(object):
    """Per-(device, stream) zero-initialised scratch shared by all BN calls (kernels leave the head zeroed):
    [2C sums | counter | pad | per-block partials]."""
    _ws = {}

    @classmethod
    def get(cls, device, C):
        lib = N.cuda()
        key = (device, torch.cuda.current_stream(device).cuda_stream)
        need = lib.MXBBnWorkspaceFloats(c_int(4096))
        ws = cls._ws.get(key)
        if ws is None or ws.numel() < need:
            ws = torch.zeros(need, dtype=torch.float32, device=device)
            cls._ws[key] = ws
        return ws


def fused_stats_workspace(x, w, force=False):
    """Workspace for GEMM-epilogue BatchNorm statistics, or None when the producer will not be our GEMM / the BN
    will not be our kernel (CPU, unsupported widths).  ``force``: the caller guarantees a GEMM producer (stem)."""
    if not ((force or _gemm_ok(x, w)) and N.have_cuda_kernels()):
        return None
    C = w.shape[0]
    vec = 8
    cv = C // vec if C % vec == 0 else 0
    if not (cv > 0 and cv <= 256 and (cv & (cv - 1)) == 0):
        return None
    return _BnWorkspace.get(x.device, C)


def bn_stats_workspace_for(device, C):
    """Like ``fused_stats_workspace`` for a producer that is known to be one of our GEMM-epilogue kernels: None when the
    BatchNorm kernel cannot consume pre-computed statistics for this width."""
    cv = C // 8 if C % 8 == 0 else 0
    if not (cv > 0 and cv <= 256 and (cv & (cv - 1)) == 0 and N.have_cuda_kernels()):
        return None
    return _BnWorkspace.get(torch.device(device), C)


def bn_stats_workspace(device, C):
    """Workspace whose first 2C floats a producer GEMM may fill with column sums / sums of squares."""
    return _BnWorkspace.get(torch.device(device), C)


def _bn_shape_ok(x2):
    C = x2.shape[1]
    vec = 16 // x2.element_size()
    cv = C // vec if C % vec == 0 else 0
    return cv > 0 and cv <= 256 and (cv & (cv - 1)) == 0 and x2.dtype in (torch.bfloat16, torch.float32)


class _BnActFn(torch.autograd.Function):
    """x: [M, C] contiguous (channels-last pixels).  Returns y = act(bn(x) + residual)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, residual, relu, eps, momentum, training, fix_gamma,
                pre_stats, sink_g, sink_b):
        lib = N.cuda()
        M, C = x.shape
        dev = x.device
        stats = torch.empty((4, C), dtype=torch.float32, device=dev)   # mean, invstd, scale, shift
        g32 = None if fix_gamma else gamma.float()
        b32 = beta.float()
        st = N.stream_of(x)
        if training:
            ws = pre_stats if pre_stats is not None else _BnWorkspace.get(dev, C)
            N.check(lib, lib.MXBBnStatsFinalize(_p(x), c_int(code(x.dtype)), c_int64(M), c_int(C), _p(ws), _p(g32),
                                                _p(b32), _p(running_mean), _p(running_var), _p(stats[0]), _p(stats[1]),
                                                _p(stats[2]), _p(stats[3]), c_float(eps), c_float(momentum),
                                                c_int(1 if pre_stats is not None else 0), st))
            LAUNCHES[0] += 1
        else:
            invstd = torch.rsqrt(running_var.float() + eps)
            gg = g32 if g32 is not None else torch.ones_like(invstd)
            stats[0] = running_mean.float()
            stats[1] = invstd
            stats[2] = gg * invstd
            stats[3] = b32 - running_mean.float() * gg * invstd
        y = torch.empty_like(x)
        vec = 16 // x.element_size()
        mask = torch.empty(M * (C // vec), dtype=torch.uint8, device=dev) if relu else None
        N.check(lib, lib.MXBBnApply(_p(x), _p(residual), _p(y), _p(mask), c_int(code(x.dtype)), c_int64(M), c_int(C),
                                    _p(stats[2]), _p(stats[3]), c_int(1 if relu else 0), st))
        LAUNCHES[0] += 1
        ctx.save_for_backward(x, mask, stats, g32)
        ctx.cfg = (relu, residual is not None, fix_gamma, gamma.dtype, beta.dtype)
        ctx.sink_g = sink_g if (not fix_gamma and _sink_ok(sink_g, gamma.shape)) else None
        ctx.sink_b = sink_b if _sink_ok(sink_b, beta.shape) else None
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = N.cuda()
        x, mask, stats, g32 = ctx.saved_tensors
        relu, has_res, fix_gamma, gdt, bdt = ctx.cfg
        sink_touch(ctx.sink_g, ctx.sink_b)
        M, C = x.shape
        if not dy.is_contiguous():
            dy = dy.contiguous()
        dx = torch.empty_like(x)
        dres = torch.empty_like(x) if has_res else None
        coef = torch.empty((3, C), dtype=torch.float32, device=x.device)
        ws = _BnWorkspace.get(x.device, C)
        direct = ctx.sink_b is not None and (fix_gamma or ctx.sink_g is not None)
        if direct:
            dg_ptr, db_ptr, acc = ctx.sink_g, ctx.sink_b, 1
            dgb = None
        else:
            dgb = torch.empty((2, C), dtype=torch.float32, device=x.device)
            dg_ptr, db_ptr, acc = dgb[0], dgb[1], 0
        N.check(lib, lib.MXBBnBackward(_p(dy), _p(x), _p(mask), _p(dx), _p(dres), c_int(code(x.dtype)), c_int64(M),
                                       c_int(C), _p(stats[0]), _p(stats[1]), _p(g32), _p(dg_ptr), _p(db_ptr), _p(ws),
                                       _p(coef), c_int(1 if relu else 0), c_int(acc), N.stream_of(x)))
        LAUNCHES[0] += 2
        if direct:
            dgamma = dbeta = None
        else:
            dgamma = torch.zeros_like(dgb[0]).to(gdt) if fix_gamma else dgb[0].to(gdt)
            dbeta = dgb[1].to(bdt)
        return dx, dgamma, dbeta, None, None, dres, None, None, None, None, None, None, None, None


def bn_act(x, gamma, beta, running_mean, running_var, residual=None, relu=False, eps=1e-5, momentum=0.9,
           training=True, fix_gamma=False, pre_stats=None, sink_g=None, sink_b=None):
    """BatchNorm over the last (channel) dim of a channels-last tensor, + optional residual add + ReLU.

    momentum follows the reference convention: running = running*momentum + batch*(1-momentum)
    (src/operator/nn/batch_norm.cu:359-362).  running stats are fp32 tensors updated in place.
    ``pre_stats``: workspace already holding sum / sum-of-squares of x (written by the producing GEMM's epilogue).
    ``sink_g`` / ``sink_b``: fp32 gradient-arena views that receive dgamma / dbeta directly.
    """
    C = x.shape[-1]
    x2 = x.reshape(-1, C)
    if _use_native(x2) and _bn_shape_ok(x2) and x2.is_contiguous():
        r2 = residual.reshape(-1, C) if residual is not None else None
        if r2 is not None and not r2.is_contiguous():
            r2 = r2.contiguous()
        y = _BnActFn.apply(x2, gamma, beta, running_mean, running_var, r2, relu, eps, momentum, training, fix_gamma,
                           pre_stats, sink_g, sink_b)
        return y.reshape(x.shape)
    if x.is_cuda and not N.have_cuda_kernels():
        raise RuntimeError("mxb200: CUDA kernels library missing on a GPU host")
    # eager reference path (CPU / unsupported widths)
    g = torch.ones_like(gamma) if fix_gamma else gamma
    xf = x2.float()
    if training:
        mean = xf.mean(0)
        var = xf.var(0, unbiased=False)
        with torch.no_grad():
            running_mean.mul_(momentum).add_(mean.detach() * (1 - momentum))
            running_var.mul_(momentum).add_(var.detach() * (1 - momentum))
    else:
        mean, var = running_mean.float(), running_var.float()
    y = (xf - mean) * torch.rsqrt(var + eps) * g.float() + beta.float()
    if residual is not None:
        y = y + residual.reshape(-1, C).float()
    if relu:
        y = torch.relu(y)
    return y.to(x.dtype).reshape(x.shape)


# --------------------------------------------------------------------------- max pooling (NHWC)
class _MaxPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, k, s, p):
        lib = N.cuda()
        n, h, w, c = x.shape
        ho, wo = (h + 2 * p - k) // s + 1, (w + 2 * p - k) // s + 1
        y = torch.empty((n, ho, wo, c), dtype=x.dtype, device=x.device)
        idx = torch.empty((n, ho, wo, c), dtype=torch.uint8, device=x.device)
        N.check(lib, lib.MXBMaxPoolNhwcFwd(_p(x), _p(y), _p(idx), c_int(code(x.dtype)), c_int(n), c_int(h), c_int(w),
                                           c_int(c), c_int(ho), c_int(wo), c_int(k), c_int(s), c_int(p), N.stream_of(x)))
        LAUNCHES[0] += 1
        ctx.save_for_backward(idx)
        ctx.geo = (n, h, w, c, ho, wo, k, s, p)
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = N.cuda()
        idx, = ctx.saved_tensors
        n, h, w, c, ho, wo, k, s, p = ctx.geo
        if not dy.is_contiguous():
            dy = dy.contiguous()
        dx = torch.empty((n, h, w, c), dtype=dy.dtype, device=dy.device)
        N.check(lib, lib.MXBMaxPoolNhwcBwd(_p(dy), _p(idx), _p(dx), c_int(code(dy.dtype)), c_int(n), c_int(h), c_int(w),
                                           c_int(c), c_int(ho), c_int(wo), c_int(k), c_int(s), c_int(p), N.stream_of(dy)))
        LAUNCHES[0] += 1
        return dx, None, None, None


def max_pool_nhwc(x, kernel=3, stride=2, pad=1):
    """Max pooling on a channels-last [N,H,W,C] tensor (reference op: Pooling, src/operator/nn/pool.cuh)."""
    vec = 16 // x.element_size()
    if _use_native(x) and x.is_contiguous() and x.dtype in (torch.bfloat16, torch.float32) and x.shape[-1] % vec == 0:
        return _MaxPoolFn.apply(x, kernel, stride, pad)
    if x.is_cuda and not N.have_cuda_kernels():
        raise RuntimeError("mxb200: CUDA kernels library missing on a GPU host")
    return F.max_pool2d(x.permute(0, 3, 1, 2), kernel, stride, pad).permute(0, 2, 3, 1)


# --------------------------------------------------------------------------- softmax / CE
def softmax_rows(x, log=False, temperature=1.0):
    if _use_native(x) and x.dim() == 2 and x.is_contiguous() and x.dtype in (torch.float32, torch.bfloat16, torch.float16):
        lib = N.cuda()
        y = torch.empty_like(x)
        N.check(lib, lib.MXBSoftmaxRows(_p(x), _p(y), c_int(code(x.dtype)), c_int64(x.shape[0]), c_int(x.shape[1]),
                                        c_int(1 if log else 0), c_float(temperature), N.stream_of(x)))
        LAUNCHES[0] += 1
        return y
    z = x / temperature if temperature != 1.0 else x
    return torch.log_softmax(z, dim=-1) if log else torch.softmax(z, dim=-1)


def _label_code(label):
    if label.dtype is torch.int64:
        return 6
    if label.dtype is torch.float32:
        return 0
    raise TypeError("labels must be int64 or float32")


class _SoftmaxCEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, label, smooth_alpha, use_ignore, ignore_label, sample_weight):
        lib = N.cuda()
        rows, cols = x.shape
        loss = torch.empty(rows, dtype=torch.float32, device=x.device)
        dx = torch.empty_like(x)
        N.check(lib, lib.MXBSoftmaxCE(_p(x), c_int(code(x.dtype)), _p(label), c_int(_label_code(label)), _p(loss), _p(dx),
                                      c_int64(rows), c_int(cols), c_float(1.0), c_int(1 if use_ignore else 0),
                                      c_float(ignore_label), c_float(smooth_alpha), _p(sample_weight), N.stream_of(x)))
        LAUNCHES[0] += 1
        ctx.save_for_backward(dx)
        return loss

    @staticmethod
    def backward(ctx, gl):
        dx, = ctx.saved_tensors
        return (dx * gl.unsqueeze(1).to(dx.dtype)), None, None, None, None, None


def softmax_cross_entropy(x, label, smooth_alpha=0.0, use_ignore=False, ignore_label=-1.0, sample_weight=None):
    """Per-row -log softmax(x)[label] (fp32).  One kernel produces the loss and the logits gradient."""
    if _use_native(x) and x.dim() == 2 and x.is_contiguous() and x.dtype in (torch.float32, torch.bfloat16, torch.float16):
        if label.dtype not in (torch.int64, torch.float32):
            label = label.long()
        return _SoftmaxCEFn.apply(x, label.contiguous(), float(smooth_alpha), use_ignore, float(ignore_label),
                                  sample_weight)
    if x.is_cuda and not N.have_cuda_kernels():
        raise RuntimeError("mxb200: CUDA kernels library missing on a GPU host")
    lp = torch.log_softmax(x.float(), dim=-1)
    li = label.long()
    nll = -lp.gather(1, li.clamp(0, x.shape[1] - 1).unsqueeze(1)).squeeze(1)
    if smooth_alpha > 0:
        K = x.shape[1]
        other = -(lp.sum(1) - lp.gather(1, li.clamp(0, K - 1).unsqueeze(1)).squeeze(1))
        nll = (1 - smooth_alpha) * nll + smooth_alpha / (K - 1) * other
    if use_ignore:
        nll = torch.where(label.float() == ignore_label, torch.zeros_like(nll), nll)
    if sample_weight is not None:
        nll = nll * sample_weight
    return nll


def softmax_output_grad(p, label, grad_scale=1.0, ignore_label=-1.0, use_ignore=False, normalization="null",
                        smooth_alpha=0.0):
    """SoftmaxOutput backward: (p - onehot)*grad_scale/norm (src/operator/softmax_output-inl.h:160-240)."""
    rows, cols = p.shape
    scale = float(grad_scale)
    valid = None
    if normalization == "batch":
        scale /= rows
    elif normalization == "valid":
        valid = (label.float() != ignore_label).sum().clamp(min=1) if use_ignore else None
        if valid is None:
            scale /= rows
    if _use_native(p) and p.is_contiguous() and p.dtype in (torch.float32, torch.bfloat16, torch.float16):
        lib = N.cuda()
        lab = label if label.dtype in (torch.int64, torch.float32) else label.float()
        g = torch.empty_like(p)
        N.check(lib, lib.MXBSoftmaxOutputGrad(_p(p), c_int(code(p.dtype)), _p(lab.contiguous()), c_int(_label_code(lab)),
                                              _p(g), c_int64(rows), c_int(cols), c_float(scale),
                                              c_int(1 if use_ignore else 0), c_float(ignore_label),
                                              c_float(smooth_alpha), N.stream_of(p)))
        LAUNCHES[0] += 1
    else:
        li = label.long().clamp(0, cols - 1)
        tgt = torch.full_like(p, smooth_alpha / (cols - 1) if smooth_alpha > 0 else 0.0)
        tgt.scatter_(1, li.unsqueeze(1), 1.0 - smooth_alpha)
        g = (p - tgt) * scale
        if use_ignore:
            g = g * (label.float() != ignore_label).unsqueeze(1).to(g.dtype)
    if valid is not None:
        g = g / valid.to(g.dtype)
    return g


def layer_norm(x, gamma, beta, eps=1e-5):
    from . import layernorm
    return layernorm.layer_norm(x, gamma, beta, eps)


# --------------------------------------------------------------------------- NCHW-logical operators on the NHWC kernels
# The generic operators (mx.nd.Convolution / BatchNorm / Activation, and therefore Gluon nn layers and symbolic graphs
# written for the reference's NCHW layout) keep their logical layout; on a B200 the tensors are carried in torch's
# channels_last memory format, whose ``permute(0, 2, 3, 1)`` is a free contiguous NHWC view -- exactly what the kernels
# above consume.  Every helper returns None when the fast path does not apply (the caller then runs the library op).
def _fastpath():
    return N.have_cuda_kernels() and getenv("MXB200_OP_FASTPATH", True)


def nhwc_view(x):
    """Contiguous NHWC view of a 4-D NCHW-logical tensor (a transposing copy only if it was not channels_last)."""
    v = x.permute(0, 2, 3, 1)
    return v if v.is_contiguous() else v.contiguous()


def conv2d_nchw(x, w, b, stride, pad, dilate, groups):
    """2-D convolution of a logical-NCHW bf16 CUDA tensor: 1x1 kernels run as tcgen05 GEMMs over pixels (bias in the
    epilogue), everything else goes to cuDNN in channels_last so the layout chain is never broken."""
    if not (x.is_cuda and x.dim() == 4 and x.dtype is torch.bfloat16 and w.dtype is torch.bfloat16 and _fastpath()):
        return None
    o, c, kh, kw = w.shape
    if (kh, kw) == (1, 1) and tuple(pad) == (0, 0) and tuple(dilate) == (1, 1) and groups == 1 \
            and stride[0] == stride[1] and c % 8 == 0 and o % 8 == 0:
        xn = nhwc_view(x)
        if stride[0] != 1:
            xn = subsample_nhwc(xn, stride[0])
        n, h, ww, _ = xn.shape
        y = linear(xn.reshape(n * h * ww, c), w.reshape(o, c), b)
        return y.reshape(n, h, ww, o).permute(0, 3, 1, 2)
    xc = x.contiguous(memory_format=torch.channels_last)
    return F.conv2d(xc, w, b.to(x.dtype) if b is not None else None, stride, pad, dilate, groups)


def batch_norm_nchw(x, gamma, beta, running_mean, running_var, eps, momentum, training, fix_gamma, residual=None,
                    relu=False):
    """BatchNorm over axis 1 of a logical-NCHW CUDA tensor (+ optional residual add and ReLU) on the fused kernels."""
    if not (x.is_cuda and x.dim() == 4 and _fastpath() and x.dtype in (torch.bfloat16, torch.float32)
            and running_mean.dtype is torch.float32 and running_var.dtype is torch.float32):
        return None
    xn = nhwc_view(x)
    if not _bn_shape_ok(xn.reshape(-1, xn.shape[-1])):
        return None
    rn = None
    if residual is not None:
        if residual.shape != x.shape or residual.dtype != x.dtype:
            return None
        rn = nhwc_view(residual)
    y = bn_act(xn, gamma, beta, running_mean, running_var, residual=rn, relu=relu, eps=eps, momentum=momentum,
               training=training, fix_gamma=fix_gamma)
    return y.permute(0, 3, 1, 2)


def conv1x1_bn_nchw(x, w, stride, gamma, beta, running_mean, running_var, eps, momentum, training, fix_gamma,
                    residual=None, relu=False):
    """1x1 convolution (no bias) + BatchNorm (+ residual) (+ ReLU) of a logical-NCHW bf16 tensor: the GEMM epilogue emits
    the per-channel sum / sum of squares, so BatchNorm needs no statistics pass of its own."""
    if not (x.is_cuda and x.dim() == 4 and x.dtype is torch.bfloat16 and w.dtype is torch.bfloat16 and _fastpath()
            and running_mean.dtype is torch.float32 and running_var.dtype is torch.float32):
        return None
    o, c = w.shape[0], w.shape[1]
    if c % 8 or o % 8:
        return None
    xn = nhwc_view(x)
    if stride != 1:
        xn = subsample_nhwc(xn, stride)
    n, h, ww, _ = xn.shape
    x2, w2 = xn.reshape(n * h * ww, c), w.reshape(o, c)
    if not (_gemm_ok(x2, w2) and _bn_shape_ok(torch.empty((1, o), dtype=x.dtype, device="meta"))):
        return None
    rn = None
    if residual is not None:
        if tuple(residual.shape) != (n, o, h, ww) or residual.dtype != x.dtype:
            return None
        rn = nhwc_view(residual)
    ws = fused_stats_workspace(x2, w2) if training else None
    y = linear(x2, w2, None, False, ws).reshape(n, h, ww, o)
    y = bn_act(y, gamma, beta, running_mean, running_var, residual=rn, relu=relu, eps=eps, momentum=momentum,
               training=training, fix_gamma=fix_gamma, pre_stats=ws)
    return y.permute(0, 3, 1, 2)


def act_any_layout(x, fn):
    """Apply an elementwise kernel wrapper ``fn`` (which wants a contiguous last dim) to ``x`` without breaking a
    channels_last layout chain."""
    if x.dim() == 4 and not x.is_contiguous() and x.is_contiguous(memory_format=torch.channels_last):
        return fn(x.permute(0, 2, 3, 1)).permute(0, 3, 1, 2)
    return fn(x)
