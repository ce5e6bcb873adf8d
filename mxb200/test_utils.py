"""Testing helpers (reference: python/mxnet/test_utils.py)."""
import numpy as np
import torch

from . import ndarray as nd
from .context import cpu, gpu, num_gpus, current_context
from .ndarray import NDArray


def default_context():
    return current_context()


def list_gpus():
    return list(range(num_gpus()))


def assert_almost_equal(a, b, rtol=1e-5, atol=1e-20, names=("a", "b")):
    a = a.asnumpy() if isinstance(a, NDArray) else np.asarray(a)
    b = b.asnumpy() if isinstance(b, NDArray) else np.asarray(b)
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol)


def almost_equal(a, b, rtol=1e-5, atol=1e-20):
    return np.allclose(a, b, rtol=rtol, atol=atol)


def same(a, b):
    return np.array_equal(a, b)


def rand_ndarray(shape, stype="default", density=None, dtype=None):
    arr = nd.array(np.random.uniform(-1, 1, shape), dtype=dtype)
    if stype != "default":
        mask = np.random.rand(shape[0]) < (density if density is not None else 0.5)
        a = arr.asnumpy()
        a[~mask] = 0
        return nd.array(a, dtype=dtype).tostype(stype)
    return arr


def rand_shape_2d(dim0=10, dim1=10):
    return np.random.randint(1, dim0 + 1), np.random.randint(1, dim1 + 1)


def check_numeric_gradient(fn, inputs, eps=1e-3, rtol=1e-2, atol=1e-3):
    """Finite-difference check of ``fn(*NDArrays) -> scalar NDArray`` (test_utils.py:790)."""
    from . import autograd
    arrs = [nd.array(x, dtype="float64") if not isinstance(x, NDArray) else x.astype("float64") for x in inputs]
    for a in arrs:
        a.attach_grad()
    with autograd.record():
        out = fn(*arrs)
    out.backward()
    for a in arrs:
        num = np.zeros(a.shape)
        flat = a._t.detach().reshape(-1)
        for i in range(flat.numel()):
            old = flat[i].item()
            flat[i] = old + eps
            fp = float(fn(*arrs).asnumpy().sum())
            flat[i] = old - eps
            fm = float(fn(*arrs).asnumpy().sum())
            flat[i] = old
            num.reshape(-1)[i] = (fp - fm) / (2 * eps)
        np.testing.assert_allclose(a.grad.asnumpy(), num, rtol=rtol, atol=atol)


def check_consistency(fn, inputs, ctx_list, dtypes=("float32",), rtol=1e-3, atol=1e-4):
    """Same function on several contexts / dtypes must agree (test_utils.py:1207)."""
    ref = None
    for ctx in ctx_list:
        for dt in dtypes:
            out = fn(*[nd.array(x, ctx=ctx, dtype=dt) for x in inputs]).asnumpy().astype(np.float64)
            if ref is None:
                ref = out
            else:
                np.testing.assert_allclose(out, ref, rtol=rtol, atol=atol)


def compare_optimizer(opt1, opt2, shape, dtype, w_stype="default", g_stype="default", rtol=1e-4, atol=1e-5):
    """One random weight/grad, one update_multi_precision step, compare weights (test_utils.py:1971-2000)."""
    w1 = nd.array(np.random.uniform(-1, 1, shape), dtype=dtype)
    g1 = nd.array(np.random.uniform(-1, 1, shape), dtype=dtype)
    w2, g2 = w1.copy(), g1.copy()
    s1 = opt1.create_state_multi_precision(0, w1)
    s2 = opt2.create_state_multi_precision(0, w2)
    opt1.update_multi_precision(0, w1, g1, s1)
    opt2.update_multi_precision(0, w2, g2, s2)
    np.testing.assert_allclose(w1.asnumpy(), w2.asnumpy(), rtol=rtol, atol=atol)
