"""mx.test_utils: the reference's testing helpers (python/mxnet/test_utils.py) behave as documented there."""
import numpy as np
import pytest

import mxb200 as mx
from mxb200 import test_utils as tu


def test_symbolic_forward_backward_and_numeric_gradient():
    x, w = mx.sym.var("x"), mx.sym.var("w")
    y = mx.sym.tanh(mx.sym.dot(x, w))
    xv, wv = np.random.randn(3, 4).astype("f") * 0.5, np.random.randn(4, 2).astype("f") * 0.5
    out = np.tanh(xv @ wv)
    tu.check_symbolic_forward(y, [xv, wv], [out], rtol=1e-4, atol=1e-5)
    og = np.random.randn(3, 2).astype("f")
    d = og * (1 - out ** 2)
    tu.check_symbolic_backward(y, {"x": xv, "w": wv}, [og], {"x": d @ wv.T, "w": xv.T @ d}, rtol=1e-3, atol=1e-4)
    # grad_req='add' accumulates onto the random initial gradient
    tu.check_symbolic_backward(y, [xv, wv], [og], [d @ wv.T, xv.T @ d], rtol=1e-3, atol=1e-4, grad_req="add")
    tu.check_numeric_gradient(y, [xv, wv], numeric_eps=1e-3, rtol=1e-2, atol=1e-3)
    with pytest.raises(AssertionError):
        tu.check_symbolic_forward(y, [xv, wv], [out + 0.1], rtol=1e-4, atol=1e-5)
    # short form on a python function
    tu.check_numeric_gradient(lambda a: (a * a).sum(), [np.random.rand(2, 3)])


def test_numeric_gradient_with_aux_states_and_grad_nodes():
    data = mx.sym.var("data")
    y = mx.sym.BatchNorm(data, fix_gamma=False, name="bn")
    loc = {"data": np.random.randn(4, 3).astype("f"), "bn_gamma": np.random.rand(3).astype("f") + 0.5,
           "bn_beta": np.random.randn(3).astype("f")}
    aux = {"bn_moving_mean": np.zeros(3, "f"), "bn_moving_var": np.ones(3, "f")}
    tu.check_numeric_gradient(y, loc, aux_states=aux, grad_nodes=["bn_gamma", "bn_beta"], numeric_eps=1e-3, rtol=2e-2, atol=2e-3)


def test_check_consistency_across_dtypes():
    sym = mx.sym.FullyConnected(mx.sym.var("data"), num_hidden=4, name="fc")
    ctx_list = [{"ctx": mx.cpu(), "data": (5, 8), "type_dict": {"data": "float64"}},
                {"ctx": mx.cpu(), "data": (5, 8), "type_dict": {"data": "float32"}}]
    gt = tu.check_consistency(sym, ctx_list, scale=0.5)
    assert "fc_output" in gt and "fc_weight" in gt
    # a deliberately different second symbol is caught
    bad = mx.sym.FullyConnected(mx.sym.var("data"), num_hidden=4, name="fc") * 1.5
    with pytest.raises(AssertionError):
        tu.check_consistency([sym, bad], ctx_list, scale=0.5)
    tu.check_consistency(lambda a: a * 2 + 1, [np.random.rand(3, 3)], [mx.cpu()], dtypes=("float32", "float64"))


def test_random_factories_and_comparisons():
    a, (data, idx) = tu.rand_sparse_ndarray((10, 4), "row_sparse", density=0.5)
    assert a.stype == "row_sparse" and data.shape[1:] == (4,) and len(idx) == data.shape[0]
    c, (cd, ci, cp) = tu.rand_sparse_ndarray((6, 5), "csr", density=0.3)
    assert c.stype == "csr" and len(cp) == 7 and len(cd) == len(ci)
    assert tu.create_sparse_array_zd((4, 3), "row_sparse", 0.0).asnumpy().sum() == 0
    assert tu.rand_ndarray((2, 3)).shape == (2, 3)
    assert len(tu.rand_shape_nd(4)) == 4 and len(tu.rand_shape_3d()) == 3
    x = np.arange(24.0).reshape(2, 3, 4)
    assert np.array_equal(tu.np_reduce(x, (0, 2), True, np.sum), x.sum(axis=(0, 2), keepdims=True))
    assert tu.same(x, x.copy()) and tu.almost_equal(x, x + 1e-9, atol=1e-6)
    with pytest.raises(AssertionError) as e:
        tu.assert_almost_equal(x, x + 1.0, rtol=1e-3, atol=1e-3)
    assert "Location of maximum error" in str(e.value)
    nanv = np.array([1.0, np.nan])
    tu.assert_almost_equal_ignore_nan(nanv, np.array([1.0, 5.0]))
    tu.assert_exception(lambda: 1 / 0, ZeroDivisionError)
    m = mx.nd.ones((2, 2))
    assert tu.same_array(m, m) and not tu.same_array(m, m.copy())
    assert np.array_equal(tu.assign_each(np.array([1.0, 2.0]), lambda v: v * 2), [2.0, 4.0])
    calls = []

    @tu.retry(3)
    def flaky():
        calls.append(1)
        assert len(calls) >= 2
    flaky()
    assert len(calls) == 2
    prev = tu.set_env_var("MXB200_TEST_UTILS_PROBE", "1", "unset")
    assert prev == "unset"
    with tu.discard_stderr():
        pass


def test_statistical_generator_checks():
    import scipy.stats as ss
    buckets, probs = tu.gen_buckets_probs_with_ppf(lambda q: ss.norm.ppf(q, 0, 1), 5)
    gen = lambda n: mx.nd.random.normal(0, 1, shape=(n,)).asnumpy()   # noqa: E731
    tu.verify_generator(gen, buckets, probs, nsamples=20000, nrepeat=3)
    assert tu.mean_check(gen, 0.0, 1.0, nsamples=20000) and tu.var_check(gen, 1.0, nsamples=20000)
    with pytest.raises(AssertionError):
        tu.verify_generator(lambda n: np.random.uniform(-1, 1, n), buckets, probs, nsamples=20000, nrepeat=3, success_rate=0.9)


def test_check_speed_and_simple_forward():
    s = mx.sym.FullyConnected(mx.sym.var("data"), num_hidden=8, name="fc")
    t = tu.check_speed(s, ctx=mx.cpu(), N=3, data=(4, 16))
    assert t > 0
    assert tu.check_speed(s, ctx=mx.cpu(), N=2, typ="forward", data=(4, 16)) > 0
    out = tu.simple_forward(mx.sym.var("a") + mx.sym.var("b"), a=np.ones((2, 2), "f"), b=np.ones((2, 2), "f"))
    assert np.array_equal(out, 2 * np.ones((2, 2)))
