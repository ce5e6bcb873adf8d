"""Second operator batch: the long tail of the reference's registry (src/operator/tensor/*, contrib/*, random/*,
quantization/*, legacy *_v1 aliases and the NNVM-internal ``_xxx_scalar`` names that serialized graphs refer to).
None of these is on a hot path; they run on torch ops and exist so that graphs / scripts written against the
reference resolve every operator name.  File:line cites point at the reference implementation of the semantics."""
import math

import numpy as np
import torch
import torch.nn.functional as F

from ..base import MXNetError, torch_dtype
from . import op as _o
from .ndarray import NDArray, _op
from .op import _REG, _out, register, _unary, _binary, _cmpf

# ----------------------------------------------------------------------------- unary math (elemwise_unary_op_*.cc)
_unary("arcsinh", torch.asinh)
_unary("arccosh", torch.acosh)
_unary("arctanh", torch.atanh)
class _Cbrt(torch.autograd.Function):
    """cube root with the analytic gradient 1 / (3 y^2) (inf at 0, like mshadow_op::cube_root_grad)."""

    @staticmethod
    def forward(ctx, t):
        y = torch.sign(t) * torch.abs(t).pow(1.0 / 3.0)
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, g):
        (y,) = ctx.saved_tensors
        return g / (3.0 * y * y)


_unary("cbrt", _Cbrt.apply)
_unary("rcbrt", lambda t: torch.sign(t) * torch.abs(t).pow(-1.0 / 3.0))
_unary("degrees", torch.rad2deg)
_unary("radians", torch.deg2rad)
_unary("fix", torch.trunc)
_unary("rint", lambda t: torch.where((t - torch.floor(t)) <= 0.5, torch.floor(t), torch.ceil(t)))   # ties toward -inf like the reference
_unary("_copy", lambda t: t.clone())


@register(name="_identity_with_attr_like_rhs")
def _identity_with_attr_like_rhs(lhs, rhs=None, out=None):
    """Identity on ``lhs`` whose shape / storage attributes are declared equal to ``rhs`` (a shape-inference constraint in
    the reference: src/operator/tensor/elemwise_unary_op_basic.cc)."""
    return _out(_op(lambda t, *_r: t, *([lhs] if rhs is None else [lhs, rhs])), out)


@register
def hard_sigmoid(data, alpha=0.2, beta=0.5, out=None):
    return _out(_op(lambda t: torch.clamp(alpha * t + beta, 0.0, 1.0), data), out)


@register
def softmin(data, axis=-1, temperature=None, out=None):
    tt = 1.0 if temperature is None else float(temperature)
    return _out(_op(lambda t: torch.softmax(-t.float() / tt, dim=axis).to(t.dtype), data), out)


@register
def nanprod(data, axis=None, keepdims=False, exclude=False, out=None):
    def g(t):
        t = torch.where(torch.isnan(t), torch.ones_like(t), t)
        if axis is None or axis == ():
            return t.prod().reshape((1,) * (t.dim() if keepdims else 1))
        axes = (axis,) if isinstance(axis, int) else tuple(axis)
        if exclude:
            axes = tuple(i for i in range(t.dim()) if i not in [a % t.dim() for a in axes])
        for a in sorted([a % t.dim() for a in axes], reverse=True):
            t = t.prod(dim=a, keepdim=bool(keepdims))
        return t.reshape(1) if t.dim() == 0 else t
    return _out(_op(g, data), out)


# ----------------------------------------------------------------------------- NNVM-internal names of scalar / binary ops
def _alias(new, old):
    if old in _REG and new not in _REG:
        _REG[new] = _REG[old]


def _scalar(name, fn):
    def f(data, scalar=0.0, out=None, **kw):
        return _out(_op(lambda t: fn(t, float(scalar)), data), out)
    f.__name__ = name
    _REG[name] = f
    return f


_scalar("_plus_scalar", lambda t, s: t + s)
_scalar("_minus_scalar", lambda t, s: t - s)
_scalar("_rminus_scalar", lambda t, s: s - t)
_scalar("_mul_scalar", lambda t, s: t * s)
_scalar("_div_scalar", lambda t, s: t / s)
_scalar("_rdiv_scalar", lambda t, s: s / t)
_scalar("_mod_scalar", lambda t, s: _o._mx_mod(t, s))
_scalar("_rmod_scalar", lambda t, s: torch.remainder(torch.full_like(t, s), t))
_scalar("_power_scalar", lambda t, s: torch.pow(t, s))
_scalar("_rpower_scalar", lambda t, s: torch.pow(torch.full_like(t, s), t))
_scalar("_maximum_scalar", lambda t, s: torch.clamp_min(t, s))
_scalar("_minimum_scalar", lambda t, s: torch.clamp_max(t, s))
_scalar("_hypot_scalar", lambda t, s: torch.hypot(t, torch.full_like(t, s)))
_scalar("_equal_scalar", lambda t, s: (t == s).to(t.dtype))
_scalar("_not_equal_scalar", lambda t, s: (t != s).to(t.dtype))
_scalar("_greater_scalar", lambda t, s: (t > s).to(t.dtype))
_scalar("_greater_equal_scalar", lambda t, s: (t >= s).to(t.dtype))
_scalar("_lesser_scalar", lambda t, s: (t < s).to(t.dtype))
_scalar("_lesser_equal_scalar", lambda t, s: (t <= s).to(t.dtype))
_scalar("_logical_and_scalar", lambda t, s: ((t != 0) & (s != 0)).to(t.dtype))
_scalar("_logical_or_scalar", lambda t, s: ((t != 0) | (s != 0)).to(t.dtype))
_scalar("_logical_xor_scalar", lambda t, s: ((t != 0) ^ (s != 0)).to(t.dtype))
_scalar("_scatter_plus_scalar", lambda t, s: t + s)
_scalar("_scatter_minus_scalar", lambda t, s: t - s)

_binary("_equal", _cmpf(torch.eq))
_binary("_not_equal", _cmpf(torch.ne))
_binary("_greater", _cmpf(torch.gt))
_binary("_greater_equal", _cmpf(torch.ge))
_binary("_lesser", _cmpf(torch.lt))
_binary("_lesser_equal", _cmpf(torch.le))
_binary("_maximum", _o._mx_maximum)
_binary("_minimum", _o._mx_minimum)
_binary("_mod", _o._mx_mod)
_binary("_power", torch.pow)
_binary("_hypot", torch.hypot)
_binary("_logical_and", lambda a, b: ((a != 0) & (b != 0)).to(a.dtype))
_binary("_logical_or", lambda a, b: ((a != 0) | (b != 0)).to(a.dtype))
_binary("_logical_xor", lambda a, b: ((a != 0) ^ (b != 0)).to(a.dtype))
_binary("_grad_add", torch.add)
_binary("_scatter_elemwise_div", torch.div)
for _n, _o_ in (("_linalg_gemm", "linalg_gemm"), ("_linalg_gemm2", "linalg_gemm2"), ("_linalg_potrf", "linalg_potrf"),
                ("_linalg_trsm", "linalg_trsm"), ("_linalg_sumlogdiag", "linalg_sumlogdiag"), ("_linalg_syrk", "linalg_syrk"),
                ("BatchNorm_v1", "BatchNorm"), ("CuDNNBatchNorm", "BatchNorm"), ("Convolution_v1", "Convolution"),
                ("Pooling_v1", "Pooling"), ("_shuffle", "shuffle"), ("_histogram", "histogram"),
                ("_ravel_multi_index", "ravel_multi_index"), ("_unravel_index", "unravel_index"),
                ("_sparse_retain", "sparse_retain"), ("_square_sum", "square_sum"), ("_CrossDeviceCopy", "identity")):
    _alias(_n, _o_)


# ----------------------------------------------------------------------------- init ops (tensor/init_op.cc)
def _ctx_dev(ctx):
    from ..context import current_context
    c = ctx if ctx is not None and not isinstance(ctx, str) else current_context()
    return c, c.torch_device


def _init(name, make):
    def f(shape=(), ctx=None, dtype="float32", out=None, **kw):
        c, dev = _ctx_dev(ctx)
        shape_ = (shape,) if isinstance(shape, int) else tuple(shape)
        return _out(NDArray(make(shape_, torch_dtype(dtype or "float32"), dev, kw), c), out)
    f.__name__ = name
    _REG[name] = f


_init("_zeros", lambda s, d, dev, kw: torch.zeros(s, dtype=d, device=dev))
_init("_zeros_without_dtype", lambda s, d, dev, kw: torch.zeros(s, dtype=d, device=dev))
_init("_ones", lambda s, d, dev, kw: torch.ones(s, dtype=d, device=dev))
_init("_full", lambda s, d, dev, kw: torch.full(s, float(kw.get("value", 0.0)), dtype=d, device=dev))


@register(name="_eye")
def _eye(N=1, M=0, k=0, ctx=None, dtype="float32", out=None):
    c, dev = _ctx_dev(ctx)
    m = int(M) if int(M) > 0 else int(N)
    t = torch.zeros(int(N), m, dtype=torch_dtype(dtype), device=dev)
    lo, hi = max(0, -int(k)), min(int(N), m - int(k))
    if hi > lo:                                  # a diagonal outside the matrix leaves it all zeros (numpy.eye)
        idx = torch.arange(lo, hi, device=dev)
        t[idx, idx + int(k)] = 1
    return _out(NDArray(t, c), out)


@register(name="_arange")
def _arange(start=0.0, stop=None, step=1.0, repeat=1, infer_range=False, ctx=None, dtype="float32", out=None):
    c, dev = _ctx_dev(ctx)
    if stop is None:
        start, stop = 0.0, start
    t = torch.arange(float(start), float(stop), float(step), device=dev).to(torch_dtype(dtype))
    if int(repeat) > 1:
        t = t.repeat_interleave(int(repeat))
    return _out(NDArray(t, c), out)


# ----------------------------------------------------------------------------- layout shuffles (tensor/matrix_op.cc)
@register
def depth_to_space(data, block_size, out=None):
    b = int(block_size)
    if b <= 0 or data.ndim != 4 or data.shape[1] % (b * b) != 0 or 0 in data.shape or data.shape[1] == 0:
        raise MXNetError("depth_to_space: block_size %d does not fit the %s input (depth must be a positive multiple of "
                         "block_size^2, no empty dimension)" % (b, str(data.shape)))

    def g(t):
        n, c, h, w = t.shape
        return t.reshape(n, b, b, c // (b * b), h, w).permute(0, 3, 4, 1, 5, 2).reshape(n, c // (b * b), h * b, w * b)
    return _out(_op(g, data), out)


@register
def space_to_depth(data, block_size, out=None):
    b = int(block_size)
    if b <= 0 or data.ndim != 4 or 0 in data.shape or data.shape[2] % b != 0 or data.shape[3] % b != 0:
        raise MXNetError("space_to_depth: block_size %d does not fit the %s input (height and width must be positive "
                         "multiples of block_size)" % (b, str(data.shape)))

    def g(t):
        n, c, h, w = t.shape
        return t.reshape(n, c, h // b, b, w // b, b).permute(0, 3, 5, 1, 2, 4).reshape(n, c * b * b, h // b, w // b)
    return _out(_op(g, data), out)


@register
def batch_take(a, indices, out=None):
    """out[i] = a[i, indices[i]] (tensor/indexing_op.cc)."""
    return _out(_op(lambda t, i: t.gather(1, i.long().reshape(-1, 1)).reshape(-1), a, indices), out)


@register
def khatri_rao(*args, **kw):
    """Column-wise Kronecker product (contrib/krprod.cc)."""
    out = kw.pop("out", None)

    def g(*ts):
        r = ts[0]
        for t in ts[1:]:
            r = (r.unsqueeze(1) * t.unsqueeze(0)).reshape(-1, r.shape[1])
        return r
    return _out(_op(g, *args), out)


@register(name="_slice_assign")
def _slice_assign(lhs, rhs, begin, end, step=None, out=None):
    def g(a, b):
        a = a.clone()
        idx = tuple(slice(bg, en, (st if st else None)) for bg, en, st in zip(begin, end, step or [None] * len(begin)))
        a[idx] = b
        return a
    return _out(_op(g, lhs, rhs), out)


@register(name="_slice_assign_scalar")
def _slice_assign_scalar(data, scalar=0.0, begin=(), end=(), step=None, out=None):
    def g(a):
        a = a.clone()
        idx = tuple(slice(bg, en, (st if st else None)) for bg, en, st in zip(begin, end, step or [None] * len(begin)))
        a[idx] = float(scalar)
        return a
    return _out(_op(g, data), out)


@register(name="_scatter_set_nd")
def _scatter_set_nd(lhs, rhs, indices, shape=None, out=None):
    def g(a, b, i):
        a = a.clone()
        a[tuple(i.long()[k] for k in range(i.shape[0]))] = b
        return a
    return _out(_op(g, lhs, rhs, indices), out)


@register
def Crop(*data, **kw):
    """Legacy Crop (src/operator/crop-inl.h): crop data[0] to h_w (or to the spatial size of data[1]) at offset / centre."""
    out = kw.pop("out", None)
    offset, h_w, center = kw.get("offset", (0, 0)), kw.get("h_w", (0, 0)), kw.get("center_crop", False)

    def g(t, *like):
        h, w = (like[0].shape[2], like[0].shape[3]) if like else (int(h_w[0]), int(h_w[1]))
        oy, ox = ((t.shape[2] - h) // 2, (t.shape[3] - w) // 2) if center else (int(offset[0]), int(offset[1]))
        return t[:, :, oy:oy + h, ox:ox + w]
    return _out(_op(g, *data), out)


# ----------------------------------------------------------------------------- linalg (tensor/la_op.cc)
def _la_float(a):
    return a if a.dtype in (torch.float32, torch.float64) else a.float()


@register
def linalg_potri(A, out=None):
    """inverse from the Cholesky factor: (L L^T)^-1."""
    return _out(_op(lambda L: torch.cholesky_inverse(_la_float(L)).to(L.dtype), A), out)


@register
def linalg_trmm(A, B, transpose=False, rightside=False, alpha=1.0, out=None):
    def g(a, b):
        a = torch.tril(a)
        a = a.transpose(-1, -2) if transpose else a
        return alpha * (b @ a if rightside else a @ b)
    return _out(_op(g, A, B), out)


@register
def linalg_gelqf(A, out=None):
    """A = L Q with Q orthonormal rows (LQ factorisation)."""
    def g(a):
        q, r = torch.linalg.qr(_la_float(a).transpose(-1, -2))
        sgn = torch.sign(torch.diagonal(r, dim1=-2, dim2=-1))
        sgn = torch.where(sgn == 0, torch.ones_like(sgn), sgn)
        q, r = q * sgn.unsqueeze(-2), r * sgn.unsqueeze(-1)
        return q.transpose(-1, -2).to(a.dtype), r.transpose(-1, -2).to(a.dtype)
    return _out(_op(g, A), out)


@register
def linalg_syevd(A, out=None):
    def g(a):
        w, v = torch.linalg.eigh(_la_float(a))
        # eigenvectors are rows of U; sign convention of the reference (la_op-inl.h syevd: LAPACK output as is) is not
        # unique -- normalise every row so that its largest-magnitude entry is positive
        u = v.transpose(-1, -2)
        idx = u.abs().argmax(-1, keepdim=True)
        sgn = torch.sign(torch.gather(u, -1, idx))
        u = u * torch.where(sgn == 0, torch.ones_like(sgn), sgn)
        return u.to(a.dtype), w.to(a.dtype)
    return _out(_op(g, A), out)


for _n in ("potri", "trmm", "gelqf", "syevd"):
    _alias("_linalg_" + _n, "linalg_" + _n)


# ----------------------------------------------------------------------------- random samplers (random/sample_op.cc, multisample_op.cc)
_RAND_DEFAULTS = {"uniform": (("low", 0.0), ("high", 1.0)), "normal": (("loc", 0.0), ("scale", 1.0)), "exponential": (("lam", 1.0),),
                  "gamma": (("alpha", 1.0), ("beta", 1.0)), "poisson": (("lam", 1.0),), "negative_binomial": (("k", 1.0), ("p", 1.0)),
                  "generalized_negative_binomial": (("mu", 1.0), ("alpha", 1.0))}


def _rand(name, draw):
    spec = _RAND_DEFAULTS[name.split("_", 2)[2]]

    def f(*params, **kw):
        shape = kw.pop("shape", None)
        dtype = kw.pop("dtype", None) or "float32"
        ctx = kw.pop("ctx", None)
        out = kw.pop("out", None)
        c, dev = _ctx_dev(ctx)
        shp = () if shape is None else ((shape,) if isinstance(shape, int) else tuple(shape))
        if params and isinstance(params[0], NDArray):     # _sample_*: one distribution per element of the parameters
            ps = [p._t.float() for p in params]
            full = tuple(ps[0].shape) + shp
            ex = [p.reshape(tuple(p.shape) + (1,) * len(shp)).expand(full) for p in ps]
            return _out(NDArray(draw(full, ex, ps[0].device).to(torch_dtype(dtype)), params[0].context), out)
        # scalar parameters: positional first, then by name, then the distribution's defaults (sample_op.h *Param structs)
        vals = list(params)
        if len(vals) > len(spec) and shape is None:        # legacy positional form: random_normal(loc, scale, shape)
            shape = vals[len(spec)]
            vals = vals[:len(spec)]
            shp = (shape,) if isinstance(shape, int) else tuple(shape)
        for pname, default in spec[len(vals):]:
            vals.append(kw.pop(pname, default))
        poison = None
        if name.endswith("normal") and float(vals[1]) < 0:
            from .ndarray import _poison
            poison = MXNetError("normal: scale (standard deviation) must be non-negative, got %s" % float(vals[1]))
            vals[1] = 0.0
        ex = [torch.full(shp or (1,), float(v), device=dev) for v in vals]
        res = NDArray(draw(shp or (1,), ex, dev).to(torch_dtype(dtype)), c)
        if poison is not None:
            _poison(res._t, poison)
        return _out(res, out)
    f.__name__ = name
    _REG[name] = f


def _neg_binomial(shape, ex, dev):
    k, p = ex
    lam = torch.distributions.Gamma(k, p / (1 - p)).sample()
    return torch.poisson(lam)


def _gen_neg_binomial(shape, ex, dev):
    mu, alpha = ex
    lam = torch.where(alpha > 0, torch.distributions.Gamma(1.0 / alpha.clamp_min(1e-12), 1.0 / (alpha.clamp_min(1e-12) * mu)).sample(), mu)
    return torch.poisson(lam)


for _pre in ("_random_", "_sample_"):
    _rand(_pre + "uniform", lambda s, ex, dev: ex[0] + (ex[1] - ex[0]) * torch.rand(s, device=dev))
    _rand(_pre + "normal", lambda s, ex, dev: ex[0] + ex[1] * torch.randn(s, device=dev))
    _rand(_pre + "exponential", lambda s, ex, dev: torch.distributions.Exponential(ex[0]).sample())
    _rand(_pre + "gamma", lambda s, ex, dev: torch.where(ex[0] > 0, torch._standard_gamma(ex[0].clamp_min(1e-30)) * ex[1], torch.zeros_like(ex[0])))   # alpha = 0: the point mass at 0
    _rand(_pre + "poisson", lambda s, ex, dev: torch.poisson(ex[0]))
    _rand(_pre + "negative_binomial", _neg_binomial)
    _rand(_pre + "generalized_negative_binomial", _gen_neg_binomial)


@register(name="_sample_multinomial")
def _sample_multinomial(data, shape=None, get_prob=False, dtype="int32", out=None):
    # the class index must be representable in the output dtype (sample_multinomial_op.h bound check)
    limit = {"uint8": 1 << 8, "int8": 1 << 7, "float16": 1 << 11, "float32": 1 << 24, "int32": 1 << 31}.get(np.dtype(dtype).name if not isinstance(dtype, str) else dtype)
    if limit is not None and data.shape[-1] > limit:
        raise MXNetError("sample_multinomial: %d classes exceed the largest integer dtype %s holds exactly (%d)" % (data.shape[-1], dtype, limit))

    def g(p):
        n = 1 if shape in (None, ()) else int(np.prod(shape))
        flat = p.reshape(-1, p.shape[-1]).float()
        if flat.shape[-1] < (1 << 24):
            idx = torch.multinomial(flat, n, replacement=True)
        else:                                   # torch.multinomial is limited to 2^24 categories: inverse-CDF sampling
            cdf = torch.cumsum(flat.double(), dim=1)
            u = torch.rand(flat.shape[0], n, device=flat.device, dtype=torch.float64) * cdf[:, -1:]
            idx = torch.searchsorted(cdf, u).clamp_(max=flat.shape[-1] - 1)
        tail = () if shape in (None, ()) else ((shape,) if isinstance(shape, int) else tuple(shape))
        res = idx.reshape(tuple(p.shape[:-1]) + tail).to(torch_dtype(dtype))
        if get_prob:
            lp = torch.log(flat.gather(1, idx)).reshape(res.shape)
            return res, lp
        return res
    return _out(_op(g, data), out)


@register(name="_sample_unique_zipfian")
def _sample_unique_zipfian(range_max, shape=None, ctx=None, out=None):
    """Per row: ``num_sampled`` DISTINCT classes from the log-uniform (Zipfian) distribution over [0, range_max), drawn with
    rejection of duplicates; also returns how many draws each row needed (random/unique_sample_op.h) -- the count sampled
    softmax uses to correct for the without-replacement sampling."""
    c, dev = _ctx_dev(ctx)
    batch, n = int(shape[0]), int(shape[1])
    if n > int(range_max):
        raise MXNetError("_sample_unique_zipfian: cannot draw %d distinct classes out of %d" % (n, int(range_max)))
    log_range = math.log(int(range_max) + 1)
    res = torch.empty((batch, n), dtype=torch.int64, device=dev)
    tries = torch.empty((batch,), dtype=torch.int64, device=dev)
    for b in range(batch):
        got = torch.empty(0, dtype=torch.int64, device=dev)
        ntry = 0
        while got.numel() < n:
            m = max(2 * (n - got.numel()), 64)
            draw = ((torch.rand(m, device=dev, dtype=torch.float64) * log_range).exp() - 1).to(torch.int64) % int(range_max)
            # keep first occurrences in draw order, count the draws consumed up to the one that completes the set
            allv = torch.cat([got, draw])
            uniq, inv = torch.unique(allv, return_inverse=True)
            first = torch.full((uniq.numel(),), allv.numel(), dtype=torch.int64, device=dev).scatter_reduce(
                0, inv, torch.arange(allv.numel(), device=dev), reduce="amin")
            order = torch.sort(first).values
            if uniq.numel() >= n:
                last = int(order[n - 1])
                ntry += last - got.numel() + 1
                got = allv[order[:n]]
            else:
                ntry += m
                got = allv[order]
        res[b], tries[b] = got, ntry
    return NDArray(res, c), NDArray(tries, c)


# ----------------------------------------------------------------------------- contrib: detection helpers
def _iou(a, b, fmt="corner"):
    if fmt == "center":
        a = torch.cat([a[..., :2] - a[..., 2:] / 2, a[..., :2] + a[..., 2:] / 2], -1)
        b = torch.cat([b[..., :2] - b[..., 2:] / 2, b[..., :2] + b[..., 2:] / 2], -1)
    lt = torch.maximum(a[..., :, None, :2], b[..., None, :, :2])
    rb = torch.minimum(a[..., :, None, 2:], b[..., None, :, 2:])
    wh = (rb - lt).clamp_min(0)
    inter = wh[..., 0] * wh[..., 1]
    aa = ((a[..., 2] - a[..., 0]) * (a[..., 3] - a[..., 1]))[..., :, None]
    ab = ((b[..., 2] - b[..., 0]) * (b[..., 3] - b[..., 1]))[..., None, :]
    return inter / (aa + ab - inter).clamp_min(1e-12)


@register(name="_contrib_box_iou")
def box_iou(lhs, rhs, format="corner", out=None):
    return _out(_op(lambda a, b: _iou(a.float(), b.float(), format).to(a.dtype), lhs, rhs), out)


@register(name="_contrib_box_nms")
def box_nms(data, overlap_thresh=0.5, valid_thresh=0.0, topk=-1, coord_start=2, score_index=1, id_index=-1,
            force_suppress=False, in_format="corner", out_format="corner", out=None):
    """Greedy NMS (contrib/bounding_box-inl.h): suppressed / invalid rows become -1, survivors sorted by score."""
    def g(t):
        shp = t.shape
        x = t.reshape(-1, shp[-2], shp[-1]).float().clone()
        res = torch.full_like(x, -1.0)
        for bi in range(x.shape[0]):
            d = x[bi]
            sc = d[:, score_index]
            keep = sc > valid_thresh
            order = torch.argsort(sc, descending=True)
            order = order[keep[order]]
            if topk > 0:
                order = order[:topk]
            boxes = d[order][:, coord_start:coord_start + 4]
            iou = _iou(boxes, boxes, in_format)
            alive = torch.ones(len(order), dtype=torch.bool, device=t.device)
            ids = d[order][:, id_index] if id_index >= 0 else None
            for i in range(len(order)):
                if not alive[i]:
                    continue
                sup = iou[i] > overlap_thresh
                if ids is not None and not force_suppress:
                    sup = sup & (ids == ids[i])
                sup[:i + 1] = False
                alive = alive & ~sup
            sel = d[order][alive]
            res[bi, :sel.shape[0]] = sel
        return res.reshape(shp).to(t.dtype)
    return _out(_op(g, data), out)


_alias("_contrib_box_non_maximum_suppression", "_contrib_box_nms")


@register(name="_contrib_bipartite_matching")
def bipartite_matching(data, is_ascend=False, threshold=0.0, topk=-1, out=None):
    """Greedy bipartite matching on a score matrix (contrib/bounding_box-inl.h): returns (row->col, col->row), -1 = unmatched."""
    def g(t):
        shp = t.shape
        x = t.reshape(-1, shp[-2], shp[-1]).float()
        rows = torch.full((x.shape[0], shp[-2]), -1.0, device=t.device)
        cols = torch.full((x.shape[0], shp[-1]), -1.0, device=t.device)
        for bi in range(x.shape[0]):
            flat = x[bi].reshape(-1)
            order = torch.argsort(flat, descending=not is_ascend)
            n = 0
            for idx in order.tolist():
                r, c = divmod(idx, shp[-1])
                v = flat[idx].item()
                if (v < threshold and not is_ascend) or (v > threshold and is_ascend):
                    break
                if rows[bi, r] < 0 and cols[bi, c] < 0:
                    rows[bi, r], cols[bi, c] = c, r
                    n += 1
                    if 0 < topk <= n:
                        break
        return rows.reshape(shp[:-1]), cols.reshape(shp[:-2] + (shp[-1],))
    return _out(_op(g, data), out)


@register(name="_contrib_MultiBoxPrior")
def MultiBoxPrior(data, sizes=(1.0,), ratios=(1.0,), clip=False, steps=(-1.0, -1.0), offsets=(0.5, 0.5), out=None):
    """SSD anchors (contrib/multibox_prior.cc): per pixel len(sizes)+len(ratios)-1 boxes, corner format, normalised."""
    def g(t):
        h, w = t.shape[2], t.shape[3]
        sy = steps[0] if steps[0] > 0 else 1.0 / h
        sx = steps[1] if steps[1] > 0 else 1.0 / w
        cy = (torch.arange(h, device=t.device).float() + offsets[0]) * sy
        cx = (torch.arange(w, device=t.device).float() + offsets[1]) * sx
        wh = [(sizes[0] * math.sqrt(r) * h / w if False else s * h / w / 1.0, s) for s in sizes for r in [ratios[0]]][:0]
        whs = []
        for s in sizes:
            whs.append((s * h / w * math.sqrt(ratios[0]) / 2, s / math.sqrt(ratios[0]) / 2))
        for r in ratios[1:]:
            whs.append((sizes[0] * h / w * math.sqrt(r) / 2, sizes[0] / math.sqrt(r) / 2))
        gy, gx = torch.meshgrid(cy, cx, indexing="ij")
        boxes = []
        for hw_, hh_ in whs:
            boxes.append(torch.stack([gx - hw_, gy - hh_, gx + hw_, gy + hh_], -1))
        b = torch.stack(boxes, 2).reshape(1, -1, 4)
        return b.clamp(0, 1) if clip else b
    return _out(_op(g, data), out)


@register(name="_contrib_quadratic")
def quadratic(data, a=0.0, b=0.0, c=0.0, out=None):
    return _out(_op(lambda t: a * t * t + b * t + c, data), out)


@register(name="_contrib_fft")
def fft(data, compute_size=128, out=None):
    """Real input [..., d] -> interleaved (re, im) [..., 2d] (contrib/fft-inl.h, cuFFT C2C on zero-imag input)."""
    def g(t):
        f = torch.fft.fft(t.float(), dim=-1)
        return torch.stack([f.real, f.imag], -1).reshape(t.shape[:-1] + (2 * t.shape[-1],)).to(t.dtype)
    return _out(_op(g, data), out)


@register(name="_contrib_ifft")
def ifft(data, compute_size=128, out=None):
    """Interleaved complex [..., 2d] -> real part of the (unnormalised, like cuFFT) inverse transform [..., d]."""
    def g(t):
        z = t.float().reshape(t.shape[:-1] + (t.shape[-1] // 2, 2))
        f = torch.fft.ifft(torch.complex(z[..., 0], z[..., 1]), dim=-1) * (t.shape[-1] // 2)
        return f.real.to(t.dtype)
    return _out(_op(g, data), out)


@register(name="_contrib_count_sketch")
def count_sketch(data, h, s, out_dim, processing_batch_size=32, out=None):
    """out[:, h[i]] += s[i] * data[:, i] (contrib/count_sketch-inl.h)."""
    def g(x, hh, ss):
        o = torch.zeros(x.shape[0], int(out_dim), dtype=x.dtype, device=x.device)
        return o.index_add(1, hh.reshape(-1).long(), x * ss.reshape(1, -1))
    return _out(_op(g, data, h, s), out)


# ----------------------------------------------------------------------------- quantization (quantization/*.cc)
@register(name="_contrib_quantize")
def quantize(data, min_range, max_range, out_type="uint8", out=None):
    def g(x, lo, hi):
        lo, hi = lo.float().reshape(()), hi.float().reshape(())
        if out_type == "uint8":
            scale = 255.0 / (hi - lo)
            q = torch.clamp(torch.round((x.float() - lo) * scale), 0, 255).to(torch.uint8)
            return q, lo.reshape(1), hi.reshape(1)
        real = torch.maximum(lo.abs(), hi.abs())
        q = torch.clamp(torch.round(x.float() * (127.0 / real)), -127, 127).to(torch.int8)
        return q, (-real).reshape(1), real.reshape(1)
    return _out(_op(g, data, min_range, max_range), out)


@register(name="_contrib_dequantize")
def dequantize(data, min_range, max_range, out_type="float32", out=None):
    def g(q, lo, hi):
        lo, hi = lo.float().reshape(()), hi.float().reshape(())
        if q.dtype == torch.uint8:
            return q.float() * ((hi - lo) / 255.0) + lo
        real = torch.maximum(lo.abs(), hi.abs())
        denom = 127.0 if q.dtype == torch.int8 else 2147483647.0
        return q.float() * (real / denom)
    return _out(_op(g, data, min_range, max_range), out)


@register(name="_contrib_requantize")
def requantize(data, min_range, max_range, min_calib_range=None, max_calib_range=None, out=None):
    """int32 -> int8 with either the calibrated or the observed range (quantization/requantize-inl.h)."""
    def g(q, lo, hi):
        real_in = torch.maximum(lo.float().abs(), hi.float().abs()).reshape(())
        x = q.float() * (real_in / 2147483647.0)
        if min_calib_range is not None and max_calib_range is not None:
            real = torch.tensor(max(abs(min_calib_range), abs(max_calib_range)), device=q.device)
        else:
            real = x.abs().max()
        o = torch.clamp(torch.round(x * (127.0 / real)), -127, 127).to(torch.int8)
        return o, (-real).reshape(1), real.reshape(1)
    return _out(_op(g, data, min_range, max_range), out)


@register(name="_contrib_quantized_flatten")
def quantized_flatten(data, min_data, max_data, out=None):
    return _out(_op(lambda q, lo, hi: (q.reshape(q.shape[0], -1), lo, hi), data, min_data, max_data), out)


# ----------------------------------------------------------------------------- image ops (image/image_random.cc)
@register(name="_image_to_tensor")
def image_to_tensor(data, out=None):
    def g(t):
        x = t.float() / 255.0
        return x.permute(2, 0, 1) if x.dim() == 3 else x.permute(0, 3, 1, 2)
    return _out(_op(g, data), out)


@register(name="_image_normalize")
def image_normalize(data, mean=(0.0,), std=(1.0,), out=None):
    def g(t):
        m = torch.tensor(mean, device=t.device, dtype=t.dtype).reshape(-1, 1, 1)
        s = torch.tensor(std, device=t.device, dtype=t.dtype).reshape(-1, 1, 1)
        return (t - m) / s
    return _out(_op(g, data), out)


# ----------------------------------------------------------------------------- Correlation (src/operator/correlation.cc, FlowNet)
@register
def Correlation(data1, data2, kernel_size=1, max_displacement=1, stride1=1, stride2=1, pad_size=0, is_multiply=True, out=None):
    def g(a, b):
        k, d, p = int(kernel_size), int(max_displacement), int(pad_size)
        a, b = F.pad(a, (p, p, p, p)), F.pad(b, (p, p, p, p))
        n, c, h, w = a.shape
        br = k // 2 + d
        oh, ow = (h - 2 * br + stride1 - 1) // stride1, (w - 2 * br + stride1 - 1) // stride1
        gr = d // stride2
        outs = []
        ys = torch.arange(oh, device=a.device) * stride1 + br
        xs = torch.arange(ow, device=a.device) * stride1 + br
        for dy in range(-gr, gr + 1):
            for dx in range(-gr, gr + 1):
                acc = 0
                for ky in range(-(k // 2), k // 2 + 1):
                    for kx in range(-(k // 2), k // 2 + 1):
                        pa = a[:, :, (ys + ky)][:, :, :, (xs + kx)]
                        pb = b[:, :, (ys + ky + dy * stride2)][:, :, :, (xs + kx + dx * stride2)]
                        if is_multiply:
                            acc = acc + (pa * pb).sum(1)
                        else:
                            df = pa - pb                       # |a - b| with the reference's tie rule: d/da = +1 at a == b
                            acc = acc + torch.where(df >= 0, df, -df).sum(1)
                outs.append(acc / (k * k * c))
        return torch.stack(outs, 1)
    return _out(_op(g, data1, data2), out)


# ----------------------------------------------------------------------------- cross-GPU BatchNorm (contrib/sync_batch_norm-inl.h)
@register(name="_contrib_SyncBatchNorm")
def SyncBatchNorm(data, gamma, beta, moving_mean, moving_var, eps=1e-3, momentum=0.9, fix_gamma=True,
                  use_global_stats=False, output_mean_var=False, ndev=1, key="", out=None):
    """BatchNorm whose statistics span all data-parallel ranks.  The reference exchanges per-GPU mean/var through host
    memory and a mutex/condvar barrier (sync_batch_norm-inl.h:79-245); here the per-channel (sum, sum of squares, count)
    are reduced on the device with one all_reduce across the process group (one process per GPU)."""
    from .. import autograd as _ag
    training = _ag.is_training() and not use_global_stats

    def g(x, gm, bt, mm, mv):
        red = [0] + list(range(2, x.dim()))
        shape = [1, -1] + [1] * (x.dim() - 2)
        gm_ = torch.ones_like(gm) if fix_gamma else gm
        if training:
            xf = x.float()
            cnt = torch.tensor([float(x.numel() // x.shape[1])], device=x.device)
            stats = torch.cat([xf.sum(red), (xf * xf).sum(red), cnt])
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                stats = _AllReduceSum.apply(stats)
            c = x.shape[1]
            mean = stats[:c] / stats[-1]
            var = (stats[c:2 * c] / stats[-1] - mean * mean).clamp_min(0)
            with torch.no_grad():
                mm.mul_(momentum).add_(mean.detach().to(mm.dtype), alpha=1 - momentum)
                mv.mul_(momentum).add_(var.detach().to(mv.dtype), alpha=1 - momentum)
        else:
            mean, var = mm.float(), mv.float()
        y = (x.float() - mean.reshape(shape)) * torch.rsqrt(var.reshape(shape) + eps) * gm_.float().reshape(shape) + bt.float().reshape(shape)
        return y.to(x.dtype)
    return _out(_op(g, data, gamma, beta, moving_mean, moving_var), out)


class _AllReduceSum(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t):
        import torch.distributed as dist
        t = t.clone()
        dist.all_reduce(t)
        return t

    @staticmethod
    def backward(ctx, g):
        import torch.distributed as dist
        g = g.clone()
        dist.all_reduce(g)
        return g


# ----------------------------------------------------------------------------- misc tensor ops
@register
def histogram(data=None, bins=10, range=None, out=None, a=None, **kw):
    data = a if data is None else data           # numpy-style keyword of the python front end: histogram(a=..., bins=...)
    """tensor/histogram.cc: returns (counts, bin_edges); ``bins`` is a count (with ``range``) or an edge array."""
    def g(t, *edges):
        x = _la_float(t).reshape(-1)
        if edges:
            e = edges[0].to(x.dtype)
            idx = torch.bucketize(x, e, right=True) - 1
            idx = torch.where(x == e[-1], torch.full_like(idx, e.numel() - 2), idx)
            ok = (idx >= 0) & (idx < e.numel() - 1)
            cnt = torch.zeros(e.numel() - 1, dtype=x.dtype, device=t.device).index_add(0, idx[ok], torch.ones_like(x[ok]))
            return cnt, e
        lo, hi = (float(range[0]), float(range[1])) if range is not None else (float(x.min()), float(x.max()))
        cnt = torch.histc(x, bins=int(bins), min=lo, max=hi)
        edges_np = np.linspace(lo, hi, int(bins) + 1)          # float64 edges rounded once, like numpy.histogram
        return cnt, torch.as_tensor(edges_np, dtype=x.dtype, device=t.device)
    if isinstance(bins, NDArray):
        return _out(_op(g, data, bins), out)
    return _out(_op(g, data), out)


@register
def ravel_multi_index(data, shape=None, out=None):
    def g(t):
        idx = t.long()
        mult = [int(np.prod(shape[i + 1:])) for i in np.arange(len(shape))]
        return sum(idx[i] * mult[i] for i in np.arange(len(shape))).to(t.dtype)
    return _out(_op(g, data), out)


@register
def unravel_index(data, shape=None, out=None):
    def g(t):
        idx = t.long()
        outs = []
        for i in np.arange(len(shape)):
            m = int(np.prod(shape[i + 1:]))
            outs.append((idx // m) % int(shape[i]))
        return torch.stack(outs, 0).to(t.dtype)
    return _out(_op(g, data), out)


@register
def square_sum(data, axis=None, keepdims=False, out=None):
    def g(t):
        return (t * t).sum() if axis is None else (t * t).sum(dim=axis, keepdim=keepdims)
    return _out(_op(g, data), out)


@register(name="_rnn_param_concat")
def rnn_param_concat(*data, **kw):
    out = kw.pop("out", None)
    dim = int(kw.get("dim", 0))
    return _out(_op(lambda *ts: torch.cat([t.reshape(-1) if dim == 0 else t for t in ts], dim), *data), out)


@register(name="_contrib_SparseEmbedding")
def SparseEmbedding(data, weight, input_dim=None, output_dim=None, dtype="float32", deterministic=False, out=None):
    """Embedding whose weight gradient is row-sparse in the reference; dense here (the fused optimizer is dense)."""
    return _out(_op(lambda i, w: F.embedding(i.long(), w), data, weight), out)


for _n, _o_ in (("_histogram", "histogram"), ("_ravel_multi_index", "ravel_multi_index"), ("_unravel_index", "unravel_index"),
                ("_square_sum", "square_sum")):
    _alias(_n, _o_)


def _deq(q, lo, hi):
    return _REG["_contrib_dequantize"](q, lo, hi)


def _requant8(x):
    """float NDArray -> (int8, min, max) with the observed symmetric range."""
    from .ndarray import array as _array
    real = float(x._t.abs().max()) or 1.0
    lo, hi = NDArray(torch.tensor([-real], device=x._t.device), x.context), NDArray(torch.tensor([real], device=x._t.device), x.context)
    return _REG["_contrib_quantize"](x, lo, hi, out_type="int8")


@register(name="_contrib_quantized_fully_connected")
def quantized_fully_connected(data, weight, bias=None, min_data=None, max_data=None, min_weight=None, max_weight=None,
                              min_bias=None, max_bias=None, num_hidden=None, no_bias=False, flatten=True, out=None):
    """int8 FullyConnected with the reference's interface (quantization/quantized_fully_connected.cc); computed by
    dequantise -> FullyConnected -> int32-range outputs are returned as float with their range."""
    x, w = _deq(data, min_data, max_data), _deq(weight, min_weight, max_weight)
    b = _deq(bias, min_bias, max_bias) if (bias is not None and not no_bias) else None
    y = _REG["FullyConnected"](x, w, b, num_hidden=num_hidden, no_bias=b is None, flatten=flatten)
    return _out(_requant8(y), out)


@register(name="_contrib_quantized_conv")
def quantized_conv(data, weight, bias=None, min_data=None, max_data=None, min_weight=None, max_weight=None, min_bias=None,
                   max_bias=None, out=None, **conv_kw):
    x, w = _deq(data, min_data, max_data), _deq(weight, min_weight, max_weight)
    no_bias = conv_kw.pop("no_bias", bias is None)
    b = _deq(bias, min_bias, max_bias) if (bias is not None and not no_bias) else None
    y = _REG["Convolution"](x, w, b, no_bias=b is None, **conv_kw)
    return _out(_requant8(y), out)


@register(name="_contrib_quantized_pooling")
def quantized_pooling(data, min_data, max_data, out=None, **pool_kw):
    y = _REG["Pooling"](_deq(data, min_data, max_data), **pool_kw)
    return _out(_requant8(y), out)


# ----------------------------------------------------------------------------- control flow (src/operator/control_flow.cc; python/mxnet/ndarray/contrib.py)
def _as_list(x):
    return list(x) if isinstance(x, (list, tuple)) else [x]


@register(name="_contrib_foreach")
def foreach(body, data, init_states, name="foreach"):
    """Scan ``body(data_slice, states) -> (outputs, new_states)`` over axis 0 (imperative: a python loop, recorded by
    autograd step by step -- the reference's imperative path does the same, ndarray/contrib.py:foreach)."""
    from . import op as O
    single_d, single_s = not isinstance(data, (list, tuple)), not isinstance(init_states, (list, tuple))
    datas, states = _as_list(data), _as_list(init_states)
    outs = None
    for i in range(datas[0].shape[0]):
        xs = [d[i] for d in datas]
        o, states = body(xs[0] if single_d else xs, states[0] if single_s else states)
        states = _as_list(states)
        o = _as_list(o)
        outs = [[v] for v in o] if outs is None else [acc + [v] for acc, v in zip(outs, o)]
    stacked = [O.stack(*acc, axis=0) for acc in (outs or [])]
    return (stacked[0] if len(stacked) == 1 else stacked), (states[0] if single_s else states)


@register(name="_contrib_while_loop")
def while_loop(cond, func, loop_vars, max_iterations=None, name="while_loop"):
    """Run ``func`` while ``cond`` holds (at most max_iterations); step outputs are stacked and zero-padded to
    max_iterations like the reference (ndarray/contrib.py:while_loop)."""
    from . import op as O
    if max_iterations is None:
        raise ValueError("max_iterations should be specified")
    lv = _as_list(loop_vars)
    steps, outs = 0, None
    while steps < max_iterations and bool(float(cond(*lv).asnumpy().reshape(-1)[0])):
        o, lv = func(*lv)
        lv, o = _as_list(lv), _as_list(o)
        outs = [[v] for v in o] if outs is None else [acc + [v] for acc, v in zip(outs, o)]
        steps += 1
    stacked = []
    for acc in (outs or []):
        s = O.stack(*acc, axis=0)
        if steps < max_iterations:
            pad = NDArray(torch.zeros((max_iterations - steps,) + tuple(s.shape[1:]), dtype=s._t.dtype, device=s._t.device), s.context)
            s = O.concat(s, pad, dim=0)
        stacked.append(s)
    return stacked, lv


@register(name="_contrib_cond")
def cond(pred, then_func, else_func, name="cond"):
    return then_func() if bool(float(pred.asnumpy().reshape(-1)[0])) else else_func()


_alias("_foreach", "_contrib_foreach")
_alias("_while_loop", "_contrib_while_loop")
_alias("_cond", "_contrib_cond")


# ----------------------------------------------------------------------------- region ops (contrib/psroi_pooling.cc, proposal.cc, multibox_*.cc)
@register(name="_contrib_PSROIPooling")
def PSROIPooling(data, rois, spatial_scale, output_dim, pooled_size, group_size=0, out=None):
    """Position-sensitive ROI average pooling (R-FCN): channel (c*g + gy)*g + gx feeds output bin (gy, gx) of class c."""
    g_ = int(group_size) or int(pooled_size)
    ps = int(pooled_size)

    def g(x, r):
        n_roi = r.shape[0]
        if x.device.type == "meta":                     # shape inference
            return torch.empty(n_roi, int(output_dim), ps, ps, dtype=x.dtype, device="meta")
        res = torch.zeros(n_roi, int(output_dim), ps, ps, dtype=torch.float32, device=x.device)
        H, W = x.shape[2], x.shape[3]
        for i in range(n_roi):
            b = int(r[i, 0])
            x1, y1 = round(float(r[i, 1])) * spatial_scale, round(float(r[i, 2])) * spatial_scale
            x2, y2 = (round(float(r[i, 3])) + 1) * spatial_scale, (round(float(r[i, 4])) + 1) * spatial_scale
            bw, bh = max(x2 - x1, 0.1) / ps, max(y2 - y1, 0.1) / ps
            for py in range(ps):
                for px in range(ps):
                    hs, he = int(math.floor(py * bh + y1)), int(math.ceil((py + 1) * bh + y1))
                    ws_, we = int(math.floor(px * bw + x1)), int(math.ceil((px + 1) * bw + x1))
                    hs, he, ws_, we = max(hs, 0), min(he, H), max(ws_, 0), min(we, W)
                    if he <= hs or we <= ws_:
                        continue
                    gy, gx = min(py * g_ // ps, g_ - 1), min(px * g_ // ps, g_ - 1)
                    ch = (torch.arange(int(output_dim), device=x.device) * g_ + gy) * g_ + gx
                    res[i, :, py, px] = x[b, ch, hs:he, ws_:we].float().mean(dim=(1, 2))
        return res.to(x.dtype)
    return _out(_op(g, data, rois), out)


def _gen_anchors(stride, scales, ratios, device):
    base = torch.tensor([0, 0, stride - 1, stride - 1], dtype=torch.float32, device=device)
    w, h = base[2] - base[0] + 1, base[3] - base[1] + 1
    cx, cy = base[0] + 0.5 * (w - 1), base[1] + 0.5 * (h - 1)
    out = []
    for r in ratios:
        ws = torch.round(torch.sqrt(w * h / r))
        hs = torch.round(ws * r)
        for s in scales:
            ww, hh = ws * s, hs * s
            out.append(torch.stack([cx - 0.5 * (ww - 1), cy - 0.5 * (hh - 1), cx + 0.5 * (ww - 1), cy + 0.5 * (hh - 1)]))
    return torch.stack(out)


def _nms_keep(boxes, scores, thresh, topn):
    order = torch.argsort(scores, descending=True)
    keep = []
    iou = _iou(boxes[order], boxes[order])
    alive = torch.ones(len(order), dtype=torch.bool, device=boxes.device)
    for i in range(len(order)):
        if not alive[i]:
            continue
        keep.append(int(order[i]))
        if len(keep) >= topn:
            break
        sup = iou[i] > thresh
        sup[:i + 1] = False
        alive &= ~sup
    return keep


@register(name="_contrib_Proposal")
def Proposal(cls_prob, bbox_pred, im_info, rpn_pre_nms_top_n=6000, rpn_post_nms_top_n=300, threshold=0.7, rpn_min_size=16,
             scales=(4, 8, 16, 32), ratios=(0.5, 1, 2), feature_stride=16, output_score=False, iou_loss=False, out=None):
    """RPN proposal layer (contrib/proposal.cc): anchors + deltas -> clip -> min-size filter -> top-N -> NMS -> pad."""
    def g(cp, bp, info):
        cp, bp, info = cp.detach(), bp.detach(), info.detach()      # not differentiable: the reference's backward writes zeros
        res, res_s = [], []
        A = len(scales) * len(ratios)
        for b in range(cp.shape[0]):
            H, W = cp.shape[2], cp.shape[3]
            anchors = _gen_anchors(feature_stride, scales, ratios, cp.device)
            sy, sx = torch.meshgrid(torch.arange(H, device=cp.device) * feature_stride,
                                    torch.arange(W, device=cp.device) * feature_stride, indexing="ij")
            shifts = torch.stack([sx, sy, sx, sy], -1).reshape(-1, 1, 4).float()
            anc = (anchors.reshape(1, A, 4) + shifts).reshape(-1, 4)
            scores = cp[b, A:].float().permute(1, 2, 0).reshape(-1)
            d = bp[b].float().permute(1, 2, 0).reshape(-1, 4)
            w, h = anc[:, 2] - anc[:, 0] + 1, anc[:, 3] - anc[:, 1] + 1
            cx, cy = anc[:, 0] + 0.5 * (w - 1), anc[:, 1] + 0.5 * (h - 1)
            pcx, pcy = d[:, 0] * w + cx, d[:, 1] * h + cy
            pw, ph = torch.exp(d[:, 2]) * w, torch.exp(d[:, 3]) * h
            box = torch.stack([pcx - 0.5 * (pw - 1), pcy - 0.5 * (ph - 1), pcx + 0.5 * (pw - 1), pcy + 0.5 * (ph - 1)], 1)
            ih, iw, sc = float(info[b, 0]), float(info[b, 1]), float(info[b, 2])
            box[:, 0::2] = box[:, 0::2].clamp(0, iw - 1)
            box[:, 1::2] = box[:, 1::2].clamp(0, ih - 1)
            ok = ((box[:, 2] - box[:, 0] + 1) >= rpn_min_size * sc) & ((box[:, 3] - box[:, 1] + 1) >= rpn_min_size * sc)
            scores = torch.where(ok, scores, torch.full_like(scores, -1.0))
            top = torch.argsort(scores, descending=True)[:rpn_pre_nms_top_n]
            keep = _nms_keep(box[top], scores[top], threshold, rpn_post_nms_top_n)
            sel = top[torch.tensor(keep, device=cp.device, dtype=torch.long)]
            if len(sel) < rpn_post_nms_top_n:      # pad by repeating (the reference pads with random picks)
                sel = torch.cat([sel, sel[torch.arange(rpn_post_nms_top_n - len(sel), device=cp.device) % max(len(sel), 1)]])
            res.append(torch.cat([torch.full((rpn_post_nms_top_n, 1), float(b), device=cp.device), box[sel]], 1))
            res_s.append(scores[sel].reshape(-1, 1))
        r = torch.cat(res, 0).to(cp.dtype)
        return (r, torch.cat(res_s, 0).to(cp.dtype)) if output_score else r
    return _out(_op(g, cls_prob, bbox_pred, im_info), out)


_alias("_contrib_MultiProposal", "_contrib_Proposal")


@register(name="_contrib_MultiBoxTarget")
def MultiBoxTarget(anchor, label, cls_pred, overlap_threshold=0.5, ignore_label=-1.0, negative_mining_ratio=-1.0,
                   negative_mining_thresh=0.5, minimum_negative_samples=0, variances=(0.1, 0.1, 0.2, 0.2), out=None):
    """SSD training targets (contrib/multibox_target.cc): (loc_target [B, 4A], loc_mask [B, 4A], cls_target [B, A])."""
    def g(anc, lab, cp):
        anc = anc.reshape(-1, 4).float()
        A, B = anc.shape[0], lab.shape[0]
        loc_t = torch.zeros(B, A * 4, device=anc.device)
        loc_m = torch.zeros(B, A * 4, device=anc.device)
        cls_t = torch.zeros(B, A, device=anc.device)
        aw, ah = anc[:, 2] - anc[:, 0], anc[:, 3] - anc[:, 1]
        ax, ay = (anc[:, 0] + anc[:, 2]) / 2, (anc[:, 1] + anc[:, 3]) / 2
        for b in range(B):
            gt = lab[b].float()
            gt = gt[gt[:, 0] >= 0]
            if gt.shape[0] == 0:
                continue
            iou = _iou(anc, gt[:, 1:5])                         # [A, G]
            match = torch.full((A,), -1, dtype=torch.long, device=anc.device)
            best_a = iou.argmax(0)                              # every ground truth grabs its best anchor
            match[best_a] = torch.arange(gt.shape[0], device=anc.device)
            best_iou, best_g = iou.max(1)
            extra = (match < 0) & (best_iou > overlap_threshold)
            match[extra] = best_g[extra]
            pos = match >= 0
            g_ = gt[match.clamp_min(0)]
            gw, gh = g_[:, 3] - g_[:, 1], g_[:, 4] - g_[:, 2]
            gx, gy = (g_[:, 1] + g_[:, 3]) / 2, (g_[:, 2] + g_[:, 4]) / 2
            t = torch.stack([(gx - ax) / aw / variances[0], (gy - ay) / ah / variances[1],
                             torch.log((gw / aw).clamp_min(1e-12)) / variances[2], torch.log((gh / ah).clamp_min(1e-12)) / variances[3]], 1)
            loc_t[b] = torch.where(pos.unsqueeze(1), t, torch.zeros_like(t)).reshape(-1)
            loc_m[b] = pos.unsqueeze(1).expand(-1, 4).float().reshape(-1)
            cls_t[b] = torch.where(pos, g_[:, 0] + 1, torch.zeros(A, device=anc.device))
            if negative_mining_ratio > 0:
                neg_cand = (~pos) & (best_iou < negative_mining_thresh)
                bg_prob = torch.softmax(cp[b].float(), 0)[0]
                n_neg = max(int(pos.sum().item() * negative_mining_ratio), int(minimum_negative_samples))
                order = torch.argsort(torch.where(neg_cand, bg_prob, torch.full_like(bg_prob, 2.0)))
                chosen = torch.zeros(A, dtype=torch.bool, device=anc.device)
                chosen[order[:min(n_neg, int(neg_cand.sum().item()))]] = True
                cls_t[b] = torch.where(pos | chosen, cls_t[b], torch.full_like(cls_t[b], float(ignore_label)))
        return loc_t, loc_m, cls_t
    return _out(_op(g, anchor, label, cls_pred), out)


@register(name="_contrib_MultiBoxDetection")
def MultiBoxDetection(cls_prob, loc_pred, anchor, clip=True, threshold=0.01, background_id=0, nms_threshold=0.5,
                      force_suppress=False, variances=(0.1, 0.1, 0.2, 0.2), nms_topk=-1, out=None):
    """SSD decoding + per-class NMS (contrib/multibox_detection.cc): rows [class_id, score, x1, y1, x2, y2], -1 = none."""
    def g(cp, lp, anc):
        anc = anc.reshape(-1, 4).float()
        B, C, A = cp.shape
        aw, ah = anc[:, 2] - anc[:, 0], anc[:, 3] - anc[:, 1]
        ax, ay = (anc[:, 0] + anc[:, 2]) / 2, (anc[:, 1] + anc[:, 3]) / 2
        res = torch.full((B, A, 6), -1.0, device=cp.device)
        for b in range(B):
            d = lp[b].float().reshape(-1, 4)
            cx, cy = d[:, 0] * variances[0] * aw + ax, d[:, 1] * variances[1] * ah + ay
            w, h = torch.exp(d[:, 2] * variances[2]) * aw / 2, torch.exp(d[:, 3] * variances[3]) * ah / 2
            box = torch.stack([cx - w, cy - h, cx + w, cy + h], 1)
            if clip:
                box = box.clamp(0, 1)
            p = cp[b].float().clone()
            p[background_id] = -1
            score, cid = p.max(0)
            valid = score > threshold
            idx = torch.nonzero(valid).reshape(-1)
            idx = idx[torch.argsort(score[idx], descending=True)]
            if nms_topk > 0:
                idx = idx[:nms_topk]
            iou = _iou(box[idx], box[idx])
            alive = torch.ones(len(idx), dtype=torch.bool, device=cp.device)
            for i in range(len(idx)):
                if not alive[i]:
                    continue
                sup = iou[i] > nms_threshold
                if not force_suppress:
                    sup = sup & (cid[idx] == cid[idx[i]])
                sup[:i + 1] = False
                alive &= ~sup
            sel = idx[alive]
            out_rows = torch.cat([(cid[sel].float() - (1 if background_id == 0 else 0)).unsqueeze(1), score[sel].unsqueeze(1), box[sel]], 1)
            res[b, :out_rows.shape[0]] = out_rows
        return res.to(cp.dtype)
    return _out(_op(g, cls_prob, loc_pred, anchor), out)


# ----------------------------------------------------------------------------- the fused flagship as ONE graph node
@register(name="_contrib_FusedResNet")
def FusedResNet(data, num_layers=50, version="v1b", classes=1000, thumbnail=False, in_channels=3, layout="NCHW", eps=1e-5,
                momentum=0.9, out=None, **params):
    """ResNet (v1 / v1b) executed by the B200 kernels as a single operator, so that symbolic ``Module.fit`` scripts (the
    reference's example/dynamic-training/train_resnet.py --mode symbolic) run the same fused path as the Gluon model:
    tcgen05 GEMMs for the 1x1 convolutions / stem / FC, fused BN(+add)+ReLU, NHWC activations.  Inputs: ``data`` (NCHW like
    the reference's iterators, or NHWC with layout='NHWC') and every weight / BatchNorm parameter / running statistic by name
    (models/resnet.py::resnet_param_spec); the symbol layer creates them as ``<node>_<name>`` variables.  Returns logits."""
    from .. import autograd as _ag
    from ..models import resnet as R
    spec = R.ResNetSpec(int(num_layers), classes=int(classes), version=str(version), in_channels=int(in_channels),
                        thumbnail=thumbnail in (True, 1, "1", "True", "true"))
    names = [(n, aux) for n, _s, _k, aux in R.resnet_param_spec(spec)]
    missing = [n for n, _ in names if n not in params]
    if missing:
        raise MXNetError("FusedResNet: missing inputs %s" % missing[:4])
    training = _ag.is_training()
    arrs = [params[n] for n, _ in names]

    def g(x, *ts):
        P = {n: t for (n, aux), t in zip(names, ts) if not aux}
        A = {n: t for (n, aux), t in zip(names, ts) if aux}
        if str(layout).upper() == "NCHW":
            x = x.permute(0, 2, 3, 1)
        x = x.contiguous()
        return R.resnet_forward(spec, P, A, x, training=training, eps=float(eps), momentum=float(momentum))
    return _out(_op(g, data, *arrs), out)


# ----------------------------------------------------------------------------- image augmentation operators (image/image_random-inl.h)
# HWC (or NHWC) uint8 / float images, as produced by the decoders.
def _img_op(name):
    def deco(f):
        _REG["_image_" + name] = f
        return f
    return deco


@_img_op("flip_left_right")
def image_flip_left_right(data, out=None):
    return _out(_op(lambda t: torch.flip(t, [t.dim() - 2]), data), out)


@_img_op("flip_top_bottom")
def image_flip_top_bottom(data, out=None):
    return _out(_op(lambda t: torch.flip(t, [t.dim() - 3]), data), out)


@_img_op("random_flip_left_right")
def image_random_flip_left_right(data, out=None):
    return image_flip_left_right(data, out=out) if np.random.rand() < 0.5 else _out(data, out)


@_img_op("random_flip_top_bottom")
def image_random_flip_top_bottom(data, out=None):
    return image_flip_top_bottom(data, out=out) if np.random.rand() < 0.5 else _out(data, out)


def _gray(t):
    w = torch.tensor([0.299, 0.587, 0.114], device=t.device, dtype=torch.float32)
    return (t.float() * w).sum(-1, keepdim=True)


def _blend(t, other, alpha):
    r = t.float() * alpha + other * (1.0 - alpha)
    return (r.clamp(0, 255) if t.dtype is torch.uint8 else r).to(t.dtype)


@_img_op("adjust_lighting")
def image_adjust_lighting(data, alpha, out=None):
    """AlexNet-style PCA lighting noise with the ImageNet eigen decomposition."""
    eigval = torch.tensor([55.46, 4.794, 1.148])
    eigvec = torch.tensor([[-0.5675, 0.7192, 0.4009], [-0.5808, -0.0045, -0.8140], [-0.5836, -0.6948, 0.4203]])

    def g(t):
        a = torch.tensor(list(alpha), dtype=torch.float32)
        rgb = (eigvec * a * eigval).sum(1).to(t.device)
        r = t.float() + rgb
        return (r.clamp(0, 255) if t.dtype is torch.uint8 else r).to(t.dtype)
    return _out(_op(g, data), out)


@_img_op("random_lighting")
def image_random_lighting(data, alpha_std=0.05, out=None):
    return image_adjust_lighting(data, alpha=tuple(np.random.normal(0, alpha_std, 3)), out=out)


def _rand_factor(lo, hi):
    return float(np.random.uniform(lo, hi))


@_img_op("random_brightness")
def image_random_brightness(data, min_factor, max_factor, out=None):
    a = _rand_factor(min_factor, max_factor)
    return _out(_op(lambda t: _blend(t, torch.zeros((), device=t.device), a), data), out)


@_img_op("random_contrast")
def image_random_contrast(data, min_factor, max_factor, out=None):
    a = _rand_factor(min_factor, max_factor)
    return _out(_op(lambda t: _blend(t, _gray(t).mean(), a), data), out)


@_img_op("random_saturation")
def image_random_saturation(data, min_factor, max_factor, out=None):
    a = _rand_factor(min_factor, max_factor)
    return _out(_op(lambda t: _blend(t, _gray(t), a), data), out)


@_img_op("random_hue")
def image_random_hue(data, min_factor, max_factor, out=None):
    """Rotate the chroma plane in YIQ space by a random angle (image_random-inl.h RandomHue)."""
    a = _rand_factor(min_factor, max_factor) - 1.0

    def g(t):
        h = a * np.pi
        u, w = np.cos(h), np.sin(h)
        tyiq = torch.tensor([[0.299, 0.587, 0.114], [0.596, -0.274, -0.321], [0.211, -0.523, 0.311]])
        ityiq = torch.tensor([[1.0, 0.956, 0.621], [1.0, -0.272, -0.647], [1.0, -1.107, 1.705]])
        rot = torch.tensor([[1.0, 0.0, 0.0], [0.0, u, -w], [0.0, w, u]], dtype=torch.float32)
        m = (ityiq @ rot @ tyiq).t().to(t.device)
        r = t.float() @ m
        return (r.clamp(0, 255) if t.dtype is torch.uint8 else r).to(t.dtype)
    return _out(_op(g, data), out)


@_img_op("random_color_jitter")
def image_random_color_jitter(data, brightness=0.0, contrast=0.0, saturation=0.0, hue=0.0, out=None):
    fns = []
    if brightness > 0:
        fns.append(lambda d: image_random_brightness(d, max(0, 1 - brightness), 1 + brightness))
    if contrast > 0:
        fns.append(lambda d: image_random_contrast(d, max(0, 1 - contrast), 1 + contrast))
    if saturation > 0:
        fns.append(lambda d: image_random_saturation(d, max(0, 1 - saturation), 1 + saturation))
    if hue > 0:
        fns.append(lambda d: image_random_hue(d, max(0, 1 - hue), 1 + hue))
    for i in np.random.permutation(len(fns)):
        data = fns[i](data)
    return _out(data, out)
