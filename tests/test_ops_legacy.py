"""Legacy operator aliases + the deformable / KL-regulariser / sparse-AdaGrad operators (reference registry parity:
src/operator/contrib/deformable_convolution.cc, deformable_psroi_pooling.cc, identity_attach_KL_sparse_reg.cc,
optimizer_op.cc:651, and the ``add_alias`` names of the elementwise families)."""
import numpy as np
import torch

import mxb200 as mx
from mxb200.ndarray import op_legacy
from mxb200.ndarray.sparse import RowSparseNDArray


def test_every_alias_resolves_and_matches_its_target():
    assert op_legacy._unresolved == []
    a = mx.nd.array(np.array([[1.0, 5.0, 3.0]], "f"))
    b = mx.nd.array(np.array([[2.0, 2.0, 3.0]], "f"))
    assert np.allclose(mx.nd._Plus(a, b).asnumpy(), [[3, 7, 6]])
    assert np.allclose(mx.nd._Minus(a, b).asnumpy(), [[-1, 3, 0]])
    assert np.allclose(mx.nd._Mul(a, b).asnumpy(), [[2, 10, 9]])
    assert np.allclose(mx.nd._Div(a, b).asnumpy(), [[0.5, 2.5, 1]])
    assert np.allclose(mx.nd._Greater_Equal(a, b).asnumpy(), [[0, 1, 1]])
    assert np.allclose(mx.nd._Maximum(a, b).asnumpy(), [[2, 5, 3]])
    assert np.allclose(mx.nd._RMinusScalar(a, scalar=1.0).asnumpy(), [[0, -4, -2]])
    assert np.allclose(mx.nd._PowerScalar(a, scalar=2.0).asnumpy(), [[1, 25, 9]])
    assert np.allclose(mx.nd.max_axis(a, axis=1).asnumpy(), [5])
    assert mx.nd.random_uniform(low=0, high=1, shape=(3, 2)).shape == (3, 2)
    s = mx.sym.var("x")
    y = mx.sym._MulScalar(s, scalar=3.0)
    assert np.allclose(y.eval(x=a)[0].asnumpy(), [[3, 15, 9]])


def test_deformable_convolution_zero_offset_is_convolution():
    rng = np.random.RandomState(0)
    x = mx.nd.array(rng.randn(2, 4, 9, 9).astype("f"))
    w = mx.nd.array(rng.randn(6, 4, 3, 3).astype("f"))
    b = mx.nd.array(rng.randn(6).astype("f"))
    y = mx.nd.contrib.DeformableConvolution(x, mx.nd.zeros((2, 18, 5, 5)), w, b, kernel=(3, 3), stride=(2, 2), pad=(1, 1),
                                            num_filter=6)
    ref = mx.nd.Convolution(x, w, b, kernel=(3, 3), stride=(2, 2), pad=(1, 1), num_filter=6)
    assert np.allclose(y.asnumpy(), ref.asnumpy(), atol=1e-4)


def test_deformable_convolution_offsets_and_gradients():
    rng = np.random.RandomState(1)
    x = mx.nd.array(rng.randn(2, 4, 9, 9).astype("f"))
    w = mx.nd.array(rng.randn(6, 4, 3, 3).astype("f"))
    kw = dict(kernel=(3, 3), stride=(2, 2), pad=(1, 1), num_filter=6, no_bias=True)

    def run(dy):
        off = np.zeros((2, 18, 5, 5), "f")
        off[:, 0::2] = dy
        return mx.nd.contrib.DeformableConvolution(x, mx.nd.array(off), w, **kw).asnumpy()
    # a half-pixel vertical offset is the mean of the two neighbouring integer offsets (bilinear), away from the border
    assert np.allclose(run(0.5)[:, :, 1:4], (run(0.0) + run(1.0))[:, :, 1:4] / 2, atol=1e-4)
    # groups + deformable groups, gradients reach data, offsets and weights
    x.attach_grad()
    off = mx.nd.array(rng.randn(2, 2 * 2 * 9, 9, 9).astype("f") * 0.5)
    off.attach_grad()
    w2 = mx.nd.array(rng.randn(8, 2, 3, 3).astype("f"))
    w2.attach_grad()
    with mx.autograd.record():
        z = mx.nd.contrib.DeformableConvolution(x, off, w2, kernel=(3, 3), pad=(1, 1), num_filter=8, num_group=2,
                                                num_deformable_group=2, no_bias=True)
        loss = (z * z).sum()
    loss.backward()
    assert z.shape == (2, 8, 9, 9)
    for g in (x.grad, off.grad, w2.grad):
        assert float(g.abs().sum().asscalar()) > 0


def test_deformable_convolution_symbolic_shapes():
    data, off = mx.sym.var("data"), mx.sym.var("offset")
    y = mx.sym.contrib.DeformableConvolution(data=data, offset=off, kernel=(3, 3), pad=(1, 1), num_filter=5, name="dc")
    assert "dc_weight" in y.list_arguments() and "dc_bias" in y.list_arguments()
    arg, out, _ = y.infer_shape(data=(2, 4, 7, 7), offset=(2, 18, 7, 7))
    shapes = dict(zip(y.list_arguments(), arg))
    assert shapes["dc_weight"] == (5, 4, 3, 3) and shapes["dc_bias"] == (5,) and out[0] == (2, 5, 7, 7)


def test_deformable_psroi_pooling_matches_loop_reference():
    rng = np.random.RandomState(2)
    OD, GS, PS, SP, H, W = 2, 3, 3, 2, 12, 12
    d = rng.randn(1, OD * GS * GS, H, W).astype("f")
    rois = np.array([[0, 1, 1, 8, 8], [0, 2, 3, 10, 9]], "f")
    tr = (rng.randn(2, 2 * 2, PS, PS) * 0.5).astype("f")
    tstd = 0.1
    got = mx.nd.contrib.DeformablePSROIPooling(mx.nd.array(d), mx.nd.array(rois), mx.nd.array(tr), spatial_scale=1.0,
                                               output_dim=OD, group_size=GS, pooled_size=PS, part_size=PS,
                                               sample_per_part=SP, trans_std=tstd).asnumpy()
    ncls = tr.shape[1] // 2
    cec = OD // ncls
    ref = np.zeros((2, OD, PS, PS), "f")

    def bil(img, x, y):
        x1, x2, y1, y2 = int(np.floor(x)), int(np.ceil(x)), int(np.floor(y)), int(np.ceil(y))
        dx, dy = x - x1, y - y1
        return ((1 - dx) * (1 - dy) * img[y1, x1] + (1 - dx) * dy * img[y2, x1] + dx * (1 - dy) * img[y1, x2]
                + dx * dy * img[y2, x2])
    for n in range(2):
        x1 = round(rois[n, 1]) - 0.5
        y1 = round(rois[n, 2]) - 0.5
        x2 = round(rois[n, 3]) + 1 - 0.5
        y2 = round(rois[n, 4]) + 1 - 0.5
        rw, rh = max(x2 - x1, 0.1), max(y2 - y1, 0.1)
        bw, bh = rw / PS, rh / PS
        for c in range(OD):
            for ph in range(PS):
                for pw in range(PS):
                    cls = c // cec
                    tx = tr[n, cls * 2, ph, pw] * tstd
                    ty = tr[n, cls * 2 + 1, ph, pw] * tstd
                    ws, hs = pw * bw + x1 + tx * rw, ph * bh + y1 + ty * rh
                    ch = (c * GS + min(ph * GS // PS, GS - 1)) * GS + min(pw * GS // PS, GS - 1)
                    tot, cnt = 0.0, 0
                    for ih in range(SP):
                        for iw in range(SP):
                            w_, h_ = ws + iw * bw / SP, hs + ih * bh / SP
                            if w_ < -0.5 or w_ > W - 0.5 or h_ < -0.5 or h_ > H - 0.5:
                                continue
                            w_, h_ = min(max(w_, 0.0), W - 1.0), min(max(h_, 0.0), H - 1.0)
                            tot += bil(d[0, ch], w_, h_)
                            cnt += 1
                    ref[n, c, ph, pw] = tot / cnt if cnt else 0.0
    assert np.allclose(got, ref, atol=1e-4)


def test_identity_attach_kl_sparse_reg():
    a = mx.nd.array(np.full((4, 5), 0.5, "f"))
    a.attach_grad()
    ma = mx.nd.ones((5,)) * 0.1
    with mx.autograd.record():
        y = mx.nd.IdentityAttachKLSparseReg(a, ma, sparseness_target=0.2, penalty=0.01, momentum=0.9)
    assert np.allclose(y.asnumpy(), a.asnumpy())
    y.backward()
    m = 0.9 * 0.1 + 0.1 * 0.5
    assert np.allclose(ma.asnumpy(), m, atol=1e-6)
    assert np.allclose(a.grad.asnumpy(), 1.0 + 0.01 * (-0.2 / m + 0.8 / (1 - m)), atol=1e-5)
    # symbolic form owns the moving average as an auxiliary state
    s = mx.sym.IdentityAttachKLSparseReg(mx.sym.var("data"), name="kl")
    assert s.list_auxiliary_states() == ["kl_moving_avg"]


def test_sparse_adagrad_touches_only_present_rows():
    w, h = mx.nd.ones((4, 3)), mx.nd.zeros((4, 3))
    g = RowSparseNDArray(torch.full((2, 3), 2.0), torch.tensor([1, 3]), (4, 3))
    mx.nd._sparse_adagrad_update(w, g, h, lr=0.1, epsilon=0.0)
    assert np.allclose(h.asnumpy()[:, 0], [0, 4, 0, 4])
    assert np.allclose(w.asnumpy()[:, 0], [1, 0.9, 1, 0.9])
