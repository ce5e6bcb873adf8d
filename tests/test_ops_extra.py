"""Long-tail operators (mxb200/ndarray/op_extra.py) against NumPy oracles, the way the reference's
tests/python/unittest/test_operator.py checks them."""
import numpy as np

import mxb200 as mx
from mxb200.ndarray.op import _REG


def test_internal_scalar_and_logical_names_resolve():
    x = mx.nd.array([1.0, 2.0, 3.0])
    assert np.allclose(_REG["_rminus_scalar"](x, scalar=5).asnumpy(), [4, 3, 2])
    assert np.allclose(_REG["_rdiv_scalar"](x, scalar=6).asnumpy(), [6, 3, 2])
    assert np.allclose(_REG["_greater_equal_scalar"](x, scalar=2).asnumpy(), [0, 1, 1])
    assert np.allclose(_REG["_logical_xor"](x, mx.nd.array([0.0, 1.0, 0.0])).asnumpy(), [1, 0, 1])
    assert np.allclose(_REG["_hypot_scalar"](mx.nd.array([3.0]), scalar=4).asnumpy(), [5])


def test_unary_tail_and_layout_ops():
    x = np.array([-2.5, -0.5, 0.5, 1.5, 2.5], dtype=np.float32)
    assert np.allclose(mx.nd.rint(mx.nd.array(x)).asnumpy(), [-3, -1, 0, 1, 2])
    assert np.allclose(mx.nd.fix(mx.nd.array(x)).asnumpy(), np.fix(x))
    assert np.allclose(mx.nd.cbrt(mx.nd.array([-8.0, 27.0])).asnumpy(), [-2, 3], atol=1e-5)
    assert np.allclose(mx.nd.hard_sigmoid(mx.nd.array([-10.0, 0.0, 10.0])).asnumpy(), [0, 0.5, 1])
    a = np.arange(32, dtype=np.float32).reshape(1, 8, 2, 2)
    d2s = mx.nd.depth_to_space(mx.nd.array(a), 2)
    assert d2s.shape == (1, 2, 4, 4)
    assert np.array_equal(mx.nd.space_to_depth(d2s, 2).asnumpy(), a)
    t = mx.nd.batch_take(mx.nd.array([[1.0, 2.0], [3.0, 4.0]]), mx.nd.array([1, 0]))
    assert np.allclose(t.asnumpy(), [2, 3])
    kr = mx.nd.khatri_rao(mx.nd.array([[1.0, 2.0], [3.0, 4.0]]), mx.nd.array([[1.0, 0.0], [0.0, 1.0], [1.0, 1.0]]))
    assert kr.shape == (6, 2) and np.allclose(kr.asnumpy()[:, 0], [1, 0, 1, 3, 0, 3])


def test_linalg_tail():
    rng = np.random.RandomState(0)
    a = rng.randn(4, 4).astype(np.float32)
    spd = a @ a.T + 4 * np.eye(4, dtype=np.float32)
    L = mx.nd.linalg.potrf(mx.nd.array(spd))
    inv = mx.nd.linalg_potri(L)
    assert np.allclose(inv.asnumpy() @ spd, np.eye(4), atol=1e-3)
    U, lam = mx.nd.linalg_syevd(mx.nd.array(spd))
    assert np.allclose(U.asnumpy().T @ np.diag(lam.asnumpy()) @ U.asnumpy(), spd, atol=1e-3)
    Q, Lq = mx.nd.linalg_gelqf(mx.nd.array(a[:2]))
    assert np.allclose(Lq.asnumpy() @ Q.asnumpy(), a[:2], atol=1e-4)
    assert np.allclose(Q.asnumpy() @ Q.asnumpy().T, np.eye(2), atol=1e-4)


def test_detection_and_signal_ops():
    iou = mx.nd.contrib.box_iou(mx.nd.array([[0, 0, 2, 2]]), mx.nd.array([[1, 1, 3, 3], [0, 0, 2, 2]])).asnumpy()
    assert np.allclose(iou, [[1.0 / 7.0, 1.0]])
    nms = mx.nd.contrib.box_nms(mx.nd.array([[[0, 0.9, 0, 0, 2, 2], [0, 0.8, 0, 0, 2, 2.1], [0, 0.7, 5, 5, 6, 6]]])).asnumpy()
    assert np.allclose(nms[0, :2, 1], [0.9, 0.7]) and np.all(nms[0, 2] == -1)
    x = np.array([[1.0, 2.0, 3.0, 4.0]], dtype=np.float32)
    f = mx.nd.contrib.fft(mx.nd.array(x)).asnumpy().reshape(1, 4, 2)
    ref = np.fft.fft(x)
    assert np.allclose(f[..., 0], ref.real, atol=1e-4) and np.allclose(f[..., 1], ref.imag, atol=1e-4)
    assert np.allclose(mx.nd.contrib.ifft(mx.nd.contrib.fft(mx.nd.array(x))).asnumpy(), 4 * x, atol=1e-4)   # unnormalised like cuFFT
    q, lo, hi = mx.nd.contrib.quantize(mx.nd.array([[-1.0, 0.0, 0.5]]), mx.nd.array([-1.0]), mx.nd.array([1.0]), out_type="int8")
    assert q.asnumpy().tolist() == [[-127, 0, 64]]
    assert np.allclose(mx.nd.contrib.dequantize(q, lo, hi).asnumpy(), [[-1, 0, 0.5]], atol=1e-2)


def test_control_flow_and_samplers():
    outs, st = mx.nd.contrib.foreach(lambda x, s: (x + s, x + s), mx.nd.array(np.arange(6).reshape(3, 2)), mx.nd.zeros((2,)))
    assert np.allclose(outs.asnumpy(), [[0, 1], [2, 4], [6, 9]]) and np.allclose(st.asnumpy(), [6, 9])
    o, lv = mx.nd.contrib.while_loop(lambda i, s: i < 3, lambda i, s: ([i * 2], [i + 1, s + i]),
                                     [mx.nd.array([0.0]), mx.nd.array([0.0])], max_iterations=5)
    assert np.allclose(o[0].asnumpy().ravel(), [0, 2, 4, 0, 0]) and float(lv[1].asnumpy()[0]) == 3.0
    assert mx.nd.contrib.cond(mx.nd.array([1.0]), lambda: 1, lambda: 2) == 1
    s = _REG["_sample_uniform"](mx.nd.array([0.0, 10.0]), mx.nd.array([1.0, 11.0]), shape=(50,)).asnumpy()
    assert s.shape == (2, 50) and (s[0] < 1).all() and (s[1] >= 10).all()
    assert _REG["_random_poisson"](4.0, shape=(1000,)).asnumpy().mean() > 3


def test_ssd_ops_shapes_and_matching():
    anc = mx.nd.contrib.MultiBoxPrior(mx.nd.zeros((1, 3, 4, 4)), sizes=(0.5, 0.25), ratios=(1, 2))
    assert anc.shape == (1, 4 * 4 * 3, 4)
    lab = mx.nd.array([[[0, 0.1, 0.1, 0.5, 0.5], [-1, 0, 0, 0, 0]]])
    lt, lm, ct = mx.nd.contrib.MultiBoxTarget(anc, lab, mx.nd.zeros((1, 2, anc.shape[1])))
    assert lt.shape == (1, 192) and lm.asnumpy().sum() >= 4 and ct.asnumpy().max() == 1.0
    det = mx.nd.contrib.MultiBoxDetection(mx.nd.array(np.random.RandomState(0).rand(1, 2, 48).astype("f")), mx.nd.zeros((1, 192)), anc)
    assert det.shape == (1, 48, 6)
