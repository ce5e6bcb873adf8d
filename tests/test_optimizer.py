"""Every optimizer against a pure-python reference implementation over an option grid -- the
``compare_optimizer`` strategy of the reference (tests/python/unittest/test_optimizer.py, test_utils.py:1971)."""
import itertools
import math

import numpy as np
import pytest
import torch

import mxb200 as mx


def _clip(g, c):
    return np.clip(g, -c, c) if c is not None else g


class PySGD:
    def __init__(self, lr, momentum=0.0, wd=0.0, rescale_grad=1.0, clip_gradient=None, multi_precision=False):
        self.__dict__.update(locals())
        self.mom, self.w32 = None, None

    def step(self, w, g):
        w32 = w.astype(np.float32) if self.w32 is None else self.w32
        g = _clip(g.astype(np.float32) * self.rescale_grad, self.clip_gradient)
        if self.momentum == 0:
            w32 = (1 - self.lr * self.wd) * w32 - self.lr * g
        else:
            if self.mom is None:
                self.mom = np.zeros_like(w32)
            self.mom = self.momentum * self.mom - self.lr * self.wd * w32 - self.lr * g
            w32 = w32 + self.mom
        if self.multi_precision:
            self.w32 = w32
        return w32.astype(w.dtype)


class PyNAG(PySGD):
    def step(self, w, g):
        w32 = w.astype(np.float32)
        g = _clip(g.astype(np.float32) * self.rescale_grad, self.clip_gradient)
        if self.momentum == 0:
            return (w32 - self.lr * (g + self.wd * w32)).astype(w.dtype)
        if self.mom is None:
            self.mom = np.zeros_like(w32)
        g = g + self.wd * w32
        self.mom = self.momentum * self.mom + g
        return (w32 - self.lr * (g + self.momentum * self.mom)).astype(w.dtype)


class PyAdam:
    def __init__(self, lr, beta1=0.9, beta2=0.999, eps=1e-8, wd=0.0, rescale_grad=1.0, clip_gradient=None):
        self.__dict__.update(locals())
        self.m = self.v = None
        self.t = 0

    def step(self, w, g):
        self.t += 1
        if self.m is None:
            self.m, self.v = np.zeros_like(w), np.zeros_like(w)
        lr = self.lr * math.sqrt(1 - self.beta2 ** self.t) / (1 - self.beta1 ** self.t)
        g = _clip(g * self.rescale_grad + self.wd * w, self.clip_gradient)
        self.m = self.beta1 * self.m + (1 - self.beta1) * g
        self.v = self.beta2 * self.v + (1 - self.beta2) * g * g
        return w - lr * self.m / (np.sqrt(self.v) + self.eps)


def _run(opt, py, dtype=np.float32, steps=3, shape=(5, 7)):
    rng = np.random.RandomState(3)
    w0 = rng.randn(*shape).astype(dtype)
    w_mx, w_py = mx.nd.array(w0, dtype=dtype), w0.copy()
    upd = mx.optimizer.get_updater(opt)
    for _ in range(steps):
        g = rng.randn(*shape).astype(dtype)
        upd(0, mx.nd.array(g, dtype=dtype), w_mx)
        w_py = py.step(w_py, g)
    rtol, atol = (1e-3, 1e-3) if dtype == np.float16 else (1e-4, 1e-5)
    np.testing.assert_allclose(w_mx.asnumpy().astype(np.float32), w_py.astype(np.float32), rtol=rtol, atol=atol)


GRID = list(itertools.product([0.0, 0.9], [None, 0.4], [1.0, 0.14], [0.0, 0.03]))


@pytest.mark.parametrize("momentum,clip,rescale,wd", GRID)
@pytest.mark.parametrize("dtype,mp", [(np.float32, False), (np.float16, False), (np.float16, True)])
def test_sgd(momentum, clip, rescale, wd, dtype, mp):
    kw = dict(momentum=momentum, wd=wd, rescale_grad=rescale, clip_gradient=clip)
    if dtype == np.float16 and not mp and momentum:
        pytest.skip("fp16 momentum without master weights is too lossy to compare tightly")
    _run(mx.optimizer.SGD(learning_rate=0.1, multi_precision=mp, **kw), PySGD(0.1, multi_precision=mp, **kw), dtype)


@pytest.mark.parametrize("momentum,clip,rescale,wd", GRID)
def test_nag(momentum, clip, rescale, wd):
    kw = dict(momentum=momentum, wd=wd, rescale_grad=rescale, clip_gradient=clip)
    _run(mx.optimizer.NAG(learning_rate=0.1, **kw), PyNAG(0.1, **kw))


@pytest.mark.parametrize("clip,rescale,wd", [(None, 1.0, 0.0), (0.5, 0.8, 0.05)])
def test_adam(clip, rescale, wd):
    kw = dict(wd=wd, rescale_grad=rescale, clip_gradient=clip)
    _run(mx.optimizer.Adam(learning_rate=0.01, **kw), PyAdam(0.01, **kw))


def test_lamb_matches_formula():
    rng = np.random.RandomState(0)
    w0, g = rng.randn(6, 4).astype(np.float32), rng.randn(6, 4).astype(np.float32)
    o = mx.optimizer.LAMB(learning_rate=0.02, wd=0.01, beta1=0.9, beta2=0.999, epsilon=1e-6)
    w = mx.nd.array(w0)
    mx.optimizer.get_updater(o)(0, mx.nd.array(g), w)
    m, v = 0.1 * g, 0.001 * g * g
    r = (m / 0.1) / (np.sqrt(v / 0.001) + 1e-6) + 0.01 * w0
    ratio = np.linalg.norm(w0) / np.linalg.norm(r)
    np.testing.assert_allclose(w.asnumpy(), w0 - 0.02 * ratio * r, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("name", ["signum", "ftml", "dcasgd", "sgld", "adagrad", "rmsprop", "adadelta", "ftrl", "adamax",
                                  "nadam", "lbsgd", "lars", "ccsgd"])
def test_every_optimizer_runs_and_moves_weights(name):
    o = mx.optimizer.create(name, learning_rate=0.01)
    w = mx.nd.array(np.ones((4, 3)))
    upd = mx.optimizer.get_updater(o)
    for _ in range(3):
        upd(0, mx.nd.array(np.full((4, 3), 0.5)), w)
    assert np.isfinite(w.asnumpy()).all() and not np.allclose(w.asnumpy(), 1.0)


def test_rmsprop_centered_and_clip_weights():
    o = mx.optimizer.RMSProp(learning_rate=0.01, centered=True, clip_weights=0.9)
    w = mx.nd.ones((3,))
    mx.optimizer.get_updater(o)(0, mx.nd.ones((3,)) * -5, w)
    assert (w.asnumpy() <= 0.9 + 1e-6).all()


def test_lr_wd_mult_and_update_count():
    o = mx.optimizer.SGD(learning_rate=1.0, wd=0.1, param_idx2name={0: "fc_weight", 1: "fc_bias"})
    assert o._get_wd(0) == pytest.approx(0.1) and o._get_wd(1) == 0.0       # bias: wd_mult 0
    o.set_lr_mult({"fc_bias": 2.0})
    assert o._get_lr(1) == 2.0
    sched = mx.lr_scheduler.FactorScheduler(step=2, factor=0.5)
    o2 = mx.optimizer.SGD(learning_rate=1.0, lr_scheduler=sched)
    upd = mx.optimizer.get_updater(o2)
    w = mx.nd.ones((1,))
    for _ in range(5):
        upd(0, mx.nd.zeros((1,)), w)
    assert o2.num_update == 5 and sched.base_lr == 0.25


def test_updater_state_roundtrip():
    o = mx.optimizer.SGD(learning_rate=0.1, momentum=0.9)
    upd = mx.optimizer.get_updater(o)
    w = mx.nd.ones((3,))
    upd(0, mx.nd.ones((3,)), w)
    blob = upd.get_states(dump_optimizer=True)
    upd2 = mx.optimizer.get_updater(mx.optimizer.SGD(learning_rate=0.5))
    upd2.set_states(blob)
    assert upd2.optimizer.lr == 0.1
    np.testing.assert_allclose(upd2.states[0].asnumpy(), upd.states[0].asnumpy())


def test_sparse_sgd_lazy_update_touches_only_present_rows():
    w = mx.nd.ones((5, 2))
    g = mx.nd.sparse.row_sparse_array((np.ones((2, 2), np.float32), [1, 3]), shape=(5, 2))
    o = mx.optimizer.SGD(learning_rate=0.5, momentum=0.9)
    mx.optimizer.get_updater(o)(0, g, w)
    out = w.asnumpy()
    assert np.allclose(out[[0, 2, 4]], 1.0) and np.allclose(out[[1, 3]], 0.5)
