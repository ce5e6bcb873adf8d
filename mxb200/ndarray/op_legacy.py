"""Legacy operator names and the last few contrib operators of the reference registry.

* CamelCase / pre-NNVM aliases (``_Plus``, ``_MulScalar``, ``_Greater_Equal``, ``max_axis``, ``random_uniform`` ...): the
  reference registers them with ``.add_alias`` (src/operator/tensor/elemwise_binary_op_basic.cc:78,
  elemwise_binary_scalar_op_basic.cc, broadcast_reduce_op_value.cc:169, random/sample_op.cc:59-160) so that symbol JSON
  written by old versions still loads; here they point at the same python callables.
* ``_contrib_DeformableConvolution`` (src/operator/contrib/deformable_convolution-inl.h, nn/deformable_im2col.cuh:225-262),
  ``_contrib_DeformablePSROIPooling`` (deformable_psroi_pooling.cu:72-152), ``IdentityAttachKLSparseReg``
  (identity_attach_KL_sparse_reg-inl.h:64-107) and ``_sparse_adagrad_update`` (optimizer_op.cc:651-680) as tensor programs:
  the deformable sampling is a batched bilinear gather followed by ONE matmul per group, so the FLOPs land on the tensor
  cores through the library GEMM instead of an im2col buffer round trip per image.
"""
import torch

from ..base import MXNetError
from .ndarray import NDArray, _op
from .op import _REG, _out, register

# ----------------------------------------------------------------------------- aliases
_ALIASES = {
    # binary elementwise (same shape)
    "_Plus": "elemwise_add", "_plus": "elemwise_add", "_add": "elemwise_add",
    "_Minus": "elemwise_sub", "_minus": "elemwise_sub", "_sub": "elemwise_sub",
    "_Mul": "elemwise_mul", "_mul": "elemwise_mul",
    "_Div": "elemwise_div", "_div": "elemwise_div",
    "_Mod": "_mod", "_Power": "_power", "_Maximum": "_maximum", "_Minimum": "_minimum", "_Hypot": "_hypot",
    "_Equal": "_equal", "_Not_Equal": "_not_equal", "_Greater": "_greater", "_Greater_Equal": "_greater_equal",
    "_Lesser": "_lesser", "_Lesser_Equal": "_lesser_equal",
    "_Logical_And": "_logical_and", "_Logical_Or": "_logical_or", "_Logical_Xor": "_logical_xor",
    # tensor (op) scalar
    "_PlusScalar": "_plus_scalar", "_MinusScalar": "_minus_scalar", "_RMinusScalar": "_rminus_scalar",
    "_MulScalar": "_mul_scalar", "_DivScalar": "_div_scalar", "_RDivScalar": "_rdiv_scalar",
    "_ModScalar": "_mod_scalar", "_RModScalar": "_rmod_scalar", "_PowerScalar": "_power_scalar",
    "_RPowerScalar": "_rpower_scalar", "_MaximumScalar": "_maximum_scalar", "_MinimumScalar": "_minimum_scalar",
    "_HypotScalar": "_hypot_scalar", "_EqualScalar": "_equal_scalar", "_NotEqualScalar": "_not_equal_scalar",
    "_GreaterScalar": "_greater_scalar", "_GreaterEqualScalar": "_greater_equal_scalar",
    "_LesserScalar": "_lesser_scalar", "_LesserEqualScalar": "_lesser_equal_scalar",
    "_LogicalAndScalar": "_logical_and_scalar", "_LogicalOrScalar": "_logical_or_scalar",
    "_LogicalXorScalar": "_logical_xor_scalar",
    # reductions / slicing
    "max_axis": "max", "min_axis": "min", "_crop_assign": "_slice_assign", "_crop_assign_scalar": "_slice_assign_scalar",
    # samplers
    "random_uniform": "_random_uniform", "random_normal": "_random_normal", "random_gamma": "_random_gamma",
    "random_exponential": "_random_exponential", "random_poisson": "_random_poisson",
    "random_negative_binomial": "_random_negative_binomial",
    "random_generalized_negative_binomial": "_random_generalized_negative_binomial",
    "sample_multinomial": "_sample_multinomial",
}


def install_aliases():
    missing = []
    for alias, target in _ALIASES.items():
        if alias in _REG:
            continue
        f = _REG.get(target)
        if f is None:
            missing.append((alias, target))
            continue
        _REG[alias] = f
    return missing


# ----------------------------------------------------------------------------- deformable convolution
def _pair(v, n=2):
    if isinstance(v, str):
        v = tuple(int(s) for s in v.strip("()[] ").split(",") if s.strip())
    if isinstance(v, int):
        return (v,) * n
    v = tuple(int(x) for x in v)
    return v if len(v) == n else (v[0],) * n


def _bilinear_zero(x, h, w):
    """x: [N, C, H, W]; h, w: [N, P] sample coordinates.  Value = 0 outside [0,H) x [0,W); inside, bilinear with the upper
    neighbour clamped to the last row / column (deformable_im2col.cuh:78-111).  Returns [N, C, P]."""
    N_, C, H, W = x.shape
    valid = (h >= 0) & (w >= 0) & (h < H) & (w < W)
    hl = torch.floor(h).clamp(0, H - 1)
    wl = torch.floor(w).clamp(0, W - 1)
    lh = torch.where(hl >= H - 1, torch.zeros_like(h), h - hl)
    lw = torch.where(wl >= W - 1, torch.zeros_like(w), w - wl)
    hh_ = (hl + 1).clamp(max=H - 1)
    wh_ = (wl + 1).clamp(max=W - 1)
    flat = x.reshape(N_, C, H * W)

    def at(hi, wi):
        idx = (hi.long() * W + wi.long()).unsqueeze(1).expand(-1, C, -1)
        return flat.gather(2, idx)
    v = (at(hl, wl) * ((1 - lh) * (1 - lw)).unsqueeze(1) + at(hl, wh_) * ((1 - lh) * lw).unsqueeze(1)
         + at(hh_, wl) * (lh * (1 - lw)).unsqueeze(1) + at(hh_, wh_) * (lh * lw).unsqueeze(1))
    return v * valid.unsqueeze(1).to(v.dtype)


@register(name="_contrib_DeformableConvolution")
def DeformableConvolution(data, offset, weight, bias=None, kernel=None, stride=(1, 1), dilate=(1, 1), pad=(0, 0),
                          num_filter=None, num_group=1, num_deformable_group=1, workspace=1024, no_bias=False,
                          layout=None, out=None):
    """2-D deformable convolution v1 (Dai et al. 2017).  ``offset``: [N, 2*G_d*kh*kw, OH, OW] with channel
    2*(i*kw+j) = dy and 2*(i*kw+j)+1 = dx of tap (i, j) in every deformable group."""
    kh, kw = _pair(kernel)
    sh, sw = _pair(stride)
    dh, dw = _pair(dilate)
    ph, pw = _pair(pad)
    G, DG = int(num_group), int(num_deformable_group)

    def g(x, off, wt, *b):
        N_, C, H, W = x.shape
        OH = (H + 2 * ph - dh * (kh - 1) - 1) // sh + 1
        OW = (W + 2 * pw - dw * (kw - 1) - 1) // sw + 1
        if off.shape[1] != 2 * DG * kh * kw or off.shape[2:] != (OH, OW):
            raise MXNetError("DeformableConvolution: offset shape %s does not match (N, %d, %d, %d)"
                             % (tuple(off.shape), 2 * DG * kh * kw, OH, OW))
        xf, of = x.float(), off.float().reshape(N_, DG, kh * kw, 2, OH, OW)
        base_h = (torch.arange(OH, device=x.device, dtype=torch.float32) * sh - ph).view(1, 1, 1, OH, 1)
        base_w = (torch.arange(OW, device=x.device, dtype=torch.float32) * sw - pw).view(1, 1, 1, 1, OW)
        tap_h = (torch.arange(kh, device=x.device, dtype=torch.float32) * dh).repeat_interleave(kw).view(1, 1, kh * kw, 1, 1)
        tap_w = (torch.arange(kw, device=x.device, dtype=torch.float32) * dw).repeat(kh).view(1, 1, kh * kw, 1, 1)
        hs = base_h + tap_h + of[:, :, :, 0]                                    # [N, DG, K, OH, OW]
        ws = base_w + tap_w + of[:, :, :, 1]
        cpg = C // DG
        cols = []
        for d in range(DG):
            v = _bilinear_zero(xf[:, d * cpg:(d + 1) * cpg], hs[:, d].reshape(N_, -1), ws[:, d].reshape(N_, -1))
            cols.append(v.reshape(N_, cpg, kh * kw, OH * OW))
        col = torch.cat(cols, 1)                                                # [N, C, K, P]
        O = wt.shape[0]
        col = col.reshape(N_, G, (C // G) * kh * kw, OH * OW)
        wg = wt.float().reshape(G, O // G, (C // G) * kh * kw)
        y = torch.einsum("gok,ngkp->ngop", wg, col).reshape(N_, O, OH, OW)
        if b and not no_bias:
            y = y + b[0].float().view(1, -1, 1, 1)
        return y.to(x.dtype)
    args = [data, offset, weight] + ([bias] if (bias is not None and not no_bias) else [])
    return _out(_op(g, *args), out)


# ----------------------------------------------------------------------------- deformable PS-ROI pooling
@register(name="_contrib_DeformablePSROIPooling")
def DeformablePSROIPooling(data, rois, trans=None, spatial_scale=1.0, output_dim=None, group_size=None, pooled_size=None,
                           part_size=0, sample_per_part=1, trans_std=0.0, no_trans=False, out=None):
    """Position-sensitive ROI pooling whose bins are shifted by learned normalised offsets ``trans``
    ([R, 2*num_classes, part, part]).  Output [R, output_dim, pooled, pooled]."""
    PS, GS, OD, SP = int(pooled_size), int(group_size), int(output_dim), int(sample_per_part)
    part = int(part_size) or PS
    no_trans = no_trans in (True, 1, "1", "True", "true") or trans is None

    def g(x, r, *t):
        N_, C, H, W = x.shape
        R = r.shape[0]
        dev = x.device
        xf, rf = x.float(), r.float()
        ncls = 1 if no_trans else t[0].shape[1] // 2
        cec = OD if no_trans else OD // ncls
        bi = rf[:, 0].long()
        x1 = torch.round(rf[:, 1]) * spatial_scale - 0.5
        y1 = torch.round(rf[:, 2]) * spatial_scale - 0.5
        x2 = (torch.round(rf[:, 3]) + 1.0) * spatial_scale - 0.5
        y2 = (torch.round(rf[:, 4]) + 1.0) * spatial_scale - 0.5
        rw = (x2 - x1).clamp(min=0.1).view(R, 1, 1, 1)
        rh = (y2 - y1).clamp(min=0.1).view(R, 1, 1, 1)
        bw, bh = rw / PS, rh / PS
        pidx = torch.arange(PS, device=dev)
        ph_ = pidx.view(1, 1, PS, 1).float()
        pw_ = pidx.view(1, 1, 1, PS).float()
        parth = torch.floor(pidx.float() / PS * part).long()
        ctop = torch.arange(OD, device=dev)
        if no_trans:
            tx = ty = torch.zeros((R, OD, PS, PS), device=dev)
        else:
            tr = t[0].float().reshape(R, ncls, 2, part, part)
            cls = (ctop // cec)
            sel = tr[:, cls]                                                    # [R, OD, 2, part, part]
            sel = sel[:, :, :, parth][:, :, :, :, parth]                        # [R, OD, 2, PS, PS]
            tx, ty = sel[:, :, 0] * trans_std, sel[:, :, 1] * trans_std
        ws0 = pw_ * bw + x1.view(R, 1, 1, 1) + tx * rw                          # [R, OD, PS, PS]
        hs0 = ph_ * bh + y1.view(R, 1, 1, 1) + ty * rh
        gw = torch.floor(pidx.float() * GS / PS).long().clamp(0, GS - 1)
        chan = (ctop.view(OD, 1, 1) * GS + gw.view(1, PS, 1)) * GS + gw.view(1, 1, PS)   # [OD, PS(h), PS(w)]
        flat = xf.reshape(N_, C * H * W)
        total = torch.zeros((R, OD, PS, PS), device=dev)
        cnt = torch.zeros((R, OD, PS, PS), device=dev)
        for ih in range(SP):
            for iw in range(SP):
                w = ws0 + iw * (bw / SP)
                h = hs0 + ih * (bh / SP)
                ok = ~((w < -0.5) | (w > W - 0.5) | (h < -0.5) | (h > H - 0.5))
                w = w.clamp(0.0, W - 1.0)
                h = h.clamp(0.0, H - 1.0)
                xa, xb = torch.floor(w).long(), torch.ceil(w).long()
                ya, yb = torch.floor(h).long(), torch.ceil(h).long()
                dx, dy = w - xa, h - ya
                base = chan.view(1, OD, PS, PS) * (H * W)

                def at(yy, xx):
                    idx = (base + yy * W + xx).reshape(R, -1)
                    return flat[bi].gather(1, idx).reshape(R, OD, PS, PS)
                v = ((1 - dx) * (1 - dy) * at(ya, xa) + (1 - dx) * dy * at(yb, xa)
                     + dx * (1 - dy) * at(ya, xb) + dx * dy * at(yb, xb))
                total = total + v * ok
                cnt = cnt + ok
        return torch.where(cnt > 0, total / cnt.clamp(min=1), torch.zeros_like(total)).to(x.dtype)
    args = [data, rois] + ([] if no_trans else [trans])
    return _out(_op(g, *args), out)


# ----------------------------------------------------------------------------- KL sparseness regulariser
class _KLSparseFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, moving_avg, penalty, target, momentum):
        ctx.save_for_backward(x)
        ctx.moving_avg, ctx.cfg = moving_avg, (penalty, target, momentum)
        return x.view_as(x)

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        penalty, target, momentum = ctx.cfg
        x2 = x.reshape(x.shape[0], -1).float()
        avg = x2.mean(0)
        ma = ctx.moving_avg
        ma.mul_(momentum).add_((1 - momentum) * avg.reshape(ma.shape).to(ma.dtype))
        m = ma.float().reshape(1, -1)
        reg = penalty * (-target / m + (1 - target) / (1 - m))
        return (dy.reshape(x2.shape).float() + reg).reshape(x.shape).to(dy.dtype), None, None, None, None


@register(name="IdentityAttachKLSparseReg")
def IdentityAttachKLSparseReg(data, moving_avg=None, sparseness_target=0.1, penalty=0.001, momentum=0.9, out=None):
    """Identity in the forward pass; the backward pass adds the gradient of a KL sparseness penalty computed from a moving
    average of the mean activation per unit (aux state ``moving_avg``, updated in the backward pass like the reference)."""
    if moving_avg is None:
        moving_avg = NDArray(torch.full((int(data._t.reshape(data.shape[0], -1).shape[1]),), float(sparseness_target),
                                        device=data._t.device), data.context)
    ma = moving_avg._t
    return _out(_op(lambda x: _KLSparseFn.apply(x, ma, float(penalty), float(sparseness_target), float(momentum)), data), out)


# ----------------------------------------------------------------------------- sparse AdaGrad
@register(name="_sparse_adagrad_update")
def sparse_adagrad_update(weight, grad, history, lr, epsilon=1e-7, wd=0.0, rescale_grad=1.0, clip_gradient=-1.0, out=None):
    """AdaGrad that only touches the rows present in a row_sparse gradient (optimizer_op-inl.h:1750-1836):
    h[r] += g[r]^2 ; w[r] -= lr * g[r] / sqrt(h[r] + eps).  Dense gradients update every row."""
    from .sparse import RowSparseNDArray
    if float(wd) != 0.0:
        raise MXNetError("_sparse_adagrad_update: wd must be 0 (same restriction as the reference)")
    out = out if out is not None else weight
    with torch.no_grad():
        if isinstance(grad, RowSparseNDArray):
            rows, g = grad.indices._t.long(), grad.data._t.float()
        else:
            rows, g = None, grad._t.float()
        g = g * rescale_grad
        if clip_gradient is not None and clip_gradient >= 0:
            g = g.clamp(-clip_gradient, clip_gradient)
        if rows is None:
            history._t.add_((g * g).to(history._t.dtype))
            out._t.copy_((weight._t.float() - lr * g / torch.sqrt(history._t.float() + epsilon)).to(weight._t.dtype))
        else:
            h = history._t.index_select(0, rows).float() + g * g
            history._t.index_copy_(0, rows, h.to(history._t.dtype))
            w = weight._t.index_select(0, rows).float() - lr * g / torch.sqrt(h + epsilon)
            if out is not weight:
                out._t.copy_(weight._t)
            out._t.index_copy_(0, rows, w.to(out._t.dtype))
    return out


from . import op as _op_mod  # noqa: E402
_REG["_sparse_adagrad_update"] = _op_mod._sparse_aware_update("_sparse_adagrad_update", 1, "always")
_unresolved = install_aliases()
