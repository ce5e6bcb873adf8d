import pytest

import mxb200 as mx


def test_factor_and_multifactor():
    s = mx.lr_scheduler.FactorScheduler(step=10, factor=0.1, base_lr=1.0, stop_factor_lr=1e-3)
    assert s(5) == 1.0 and s(11) == pytest.approx(0.1) and s(1000) == pytest.approx(1e-3)
    m = mx.lr_scheduler.MultiFactorScheduler(step=[5, 10], factor=0.5, base_lr=2.0)
    assert m(5) == 2.0 and m(6) == 1.0 and m(11) == 0.5
    with pytest.raises(ValueError):
        mx.lr_scheduler.MultiFactorScheduler(step=[5, 5], factor=0.5)


def test_poly_cosine_warmup():
    p = mx.lr_scheduler.PolyScheduler(max_update=100, base_lr=1.0, pwr=2, warmup_steps=10, warmup_begin_lr=0.0)
    assert p(0) == 0 and p(5) == pytest.approx(0.5) and p(10) == pytest.approx(1.0) and p(100) == pytest.approx(0.0)
    c = mx.lr_scheduler.CosineScheduler(max_update=100, base_lr=1.0, final_lr=0.1)
    assert c(0) == pytest.approx(1.0) and c(50) == pytest.approx(0.55) and c(100) == pytest.approx(0.1)


def test_dynamic_minibatch_rule():
    base = mx.lr_scheduler.FactorScheduler(step=10 ** 6, factor=1.0, base_lr=0.4)
    d = mx.lr_scheduler.DynamicMiniBatchScheduler(base, base_batch=1024, mode="linear", smooth_steps=4)
    d.base_lr = 0.4
    assert d(1) == pytest.approx(0.4)
    d.on_membership_change(2048, num_update=10)      # 4 -> 8 workers at constant per-worker batch
    assert d(10) == pytest.approx(0.4) and d(12) == pytest.approx(0.6) and d(14) == pytest.approx(0.8)
    assert d(100) == pytest.approx(0.8)
    d.on_membership_change(1024, num_update=100)
    assert d.momentum_correction() == pytest.approx(0.5)
    assert d(104) == pytest.approx(0.4)
    # the reference's own behaviour: constant global batch -> lr untouched
    k = mx.lr_scheduler.DynamicMiniBatchScheduler(base, base_batch=1024, mode="constant_global_batch")
    k.base_lr = 0.4
    k.on_membership_change(4096)
    assert k(5) == pytest.approx(0.4)
