"""Gluon: Parameter / Block / Trainer / layers / losses (reference tests: test_gluon.py, test_gluon_trainer.py)."""
import numpy as np
import pytest

import mxb200 as mx
from mxb200 import gluon, autograd
from mxb200.gluon import nn


def test_parameter_multi_ctx_and_reset_ctx():
    p = gluon.Parameter("weight", shape=(3, 4))
    p.initialize(init="xavier", ctx=[mx.cpu(0), mx.cpu(1)])
    assert len(p.list_data()) == 2 and len(p.list_grad()) == 2
    assert p.data(mx.cpu(1)).context == mx.cpu(1) and p.data(mx.cpu(0)).shape == (3, 4)
    with pytest.raises(RuntimeError):
        p.data(mx.cpu(2))
    p.list_data()[0][:] = 1
    p.list_data()[1][:] = 3
    p.reset_ctx([mx.cpu(2)])
    assert p.list_ctx() == [mx.cpu(2)]
    np.testing.assert_allclose(p.data().asnumpy(), 2.0)      # mean of the old copies


def test_deferred_init_and_structural_save_load(tmp_path):
    net = nn.HybridSequential()
    with net.name_scope():
        net.add(nn.Dense(8, activation="relu"), nn.BatchNorm(), nn.Dense(3))
    net.initialize()
    with pytest.raises(gluon.DeferredInitializationError):
        net[0].weight.data()
    y = net(mx.nd.ones((2, 5)))
    assert y.shape == (2, 3) and net[0].weight.shape == (8, 5)
    f = str(tmp_path / "net.params")
    net.save_parameters(f)
    loaded = mx.nd.load(f)
    assert "0.weight" in loaded and "1.running_mean" in loaded       # structural names, no prefix
    net2 = nn.HybridSequential()
    with net2.name_scope():
        net2.add(nn.Dense(8, activation="relu"), nn.BatchNorm(), nn.Dense(3))
    net2.load_parameters(f)
    np.testing.assert_allclose(net2(mx.nd.ones((2, 5))).asnumpy(), net(mx.nd.ones((2, 5))).asnumpy(), rtol=1e-5)
    with pytest.raises(AssertionError):
        nn.Dense(2, in_units=5).load_parameters(f)


def test_trainer_step_matches_manual_sgd():
    x = gluon.Parameter("x", shape=(10,))
    x.initialize(ctx=[mx.cpu(0), mx.cpu(1)], init="zeros")
    trainer = gluon.Trainer([x], "sgd", {"learning_rate": 1.0, "momentum": 0.5})
    with autograd.record():
        for w in x.list_data():
            y = w + 1
            y.backward()
    trainer.step(1)
    assert (x.data(mx.cpu(1)).asnumpy() == -2).all()
    x.lr_mult = 0.5
    with autograd.record():
        for w in x.list_data():
            y = w + 1
            y.backward()
    trainer.step(1)
    assert (x.data(mx.cpu(1)).asnumpy() == -4).all()


def test_trainer_allreduce_update_split_and_states(tmp_path):
    x = gluon.Parameter("x", shape=(4,))
    x.initialize(ctx=[mx.cpu(0), mx.cpu(1)], init="zeros")
    trainer = gluon.Trainer([x], "sgd", {"learning_rate": 0.5, "momentum": 0.9}, update_on_kvstore=False)
    with autograd.record():
        for w in x.list_data():
            (w * 2).sum().backward()
    trainer.allreduce_grads()
    assert (x.list_grad()[0].asnumpy() == 4).all()
    trainer.update(1)
    np.testing.assert_allclose(x.data(mx.cpu(0)).asnumpy(), -2.0)
    f = str(tmp_path / "t.states")
    trainer.save_states(f)
    trainer.load_states(f)
    assert trainer._optimizer.momentum == 0.9


def test_trainer_reset_kv():
    for kv in ["local", "device"]:
        x = gluon.Parameter("x", shape=(3,))
        x.initialize(ctx=[mx.cpu(0), mx.cpu(1)], init="zeros")
        trainer = gluon.Trainer([x], "sgd", {"learning_rate": 0.1}, kvstore=kv)
        with autograd.record():
            for w in x.list_data():
                (w + 1).sum().backward()
        trainer.step(1)
        assert trainer._kvstore.type == kv
        x.set_data(mx.nd.ones((3,)))      # drops the kvstore: parameters re-initialised on the next step
        assert trainer._kvstore is None and not trainer._kv_initialized
        with autograd.record():
            for w in x.list_data():
                (w + 1).sum().backward()
        trainer.step(1)
        np.testing.assert_allclose(x.data(mx.cpu(0)).asnumpy(), 0.8, rtol=1e-6)


def test_losses_against_numpy():
    pred = mx.nd.array(np.random.randn(6, 5))
    label = mx.nd.array([0, 1, 2, 3, 4, 0])
    l = gluon.loss.SoftmaxCrossEntropyLoss()(pred, label).asnumpy()
    p = pred.asnumpy()
    lse = np.log(np.exp(p - p.max(1, keepdims=True)).sum(1)) + p.max(1)
    np.testing.assert_allclose(l, lse - p[np.arange(6), label.asnumpy().astype(int)], rtol=1e-5)
    a, b = mx.nd.array(np.random.randn(4, 3)), mx.nd.array(np.random.randn(4, 3))
    np.testing.assert_allclose(gluon.loss.L2Loss()(a, b).asnumpy(), ((a.asnumpy() - b.asnumpy()) ** 2).mean(1) / 2, rtol=1e-5)
    np.testing.assert_allclose(gluon.loss.L1Loss()(a, b).asnumpy(), np.abs(a.asnumpy() - b.asnumpy()).mean(1), rtol=1e-5)
    z = mx.nd.array(np.random.randn(4, 1))
    t = mx.nd.array([[1.0], [0.0], [1.0], [0.0]])
    bce = gluon.loss.SigmoidBCELoss()(z, t).asnumpy()
    zz = z.asnumpy()
    np.testing.assert_allclose(bce, (np.maximum(zz, 0) - zz * t.asnumpy() + np.log1p(np.exp(-np.abs(zz)))).mean(1), rtol=1e-5)


def test_layers_shapes_and_batchnorm_running_stats():
    x = mx.nd.array(np.random.randn(4, 3, 8, 8))
    for layer, shape in [(nn.Conv2D(6, 3, padding=1), (4, 6, 8, 8)), (nn.Conv2D(6, 3, strides=2), (4, 6, 3, 3)),
                         (nn.MaxPool2D(), (4, 3, 4, 4)), (nn.GlobalAvgPool2D(), (4, 3, 1, 1)),
                         (nn.Conv2DTranspose(5, 4, strides=2, padding=1), (4, 5, 16, 16)), (nn.Flatten(), (4, 192))]:
        layer.initialize()
        assert layer(x).shape == shape
    bn = nn.BatchNorm(momentum=0.9)
    bn.initialize()
    with autograd.record():
        y = bn(x)
    np.testing.assert_allclose(y.asnumpy().mean((0, 2, 3)), 0, atol=1e-4)
    xm = x.asnumpy().mean((0, 2, 3))
    np.testing.assert_allclose(bn.running_mean.data().asnumpy(), 0.1 * xm, rtol=1e-4, atol=1e-5)
    y_eval = bn(x)          # predict mode uses the running statistics
    assert not np.allclose(y_eval.asnumpy(), y.asnumpy())
    ln = nn.LayerNorm()
    ln.initialize()
    assert abs(float(ln(mx.nd.array(np.random.randn(2, 7))).asnumpy().mean())) < 1e-5
    emb = nn.Embedding(10, 4)
    emb.initialize()
    assert emb(mx.nd.array([1, 2, 3])).shape == (3, 4)


def test_nhwc_conv_layer_matches_nchw():
    c1 = nn.Conv2D(5, 3, padding=1, in_channels=3)
    c1.initialize()
    c2 = nn.Conv2D(5, 3, padding=1, in_channels=3, layout="NHWC", params=c1.collect_params())
    x = mx.nd.array(np.random.randn(2, 3, 6, 6))
    y1 = c1(x).asnumpy()
    y2 = c2(x.transpose(0, 2, 3, 1)).asnumpy()
    np.testing.assert_allclose(y1, y2.transpose(0, 3, 1, 2), rtol=1e-4, atol=1e-5)


def test_model_zoo_builds_and_runs():
    from mxb200.gluon.model_zoo import vision
    net = vision.get_model("resnet18_v1", classes=7)
    net.initialize()
    assert net(mx.nd.array(np.random.randn(1, 3, 32, 32))).shape == (1, 7)
    net = vision.resnet18_v2(classes=5, thumbnail=True)
    net.initialize()
    assert net(mx.nd.array(np.random.randn(2, 3, 16, 16))).shape == (2, 5)
    names = list(vision.resnet50_v1(classes=10).collect_params().keys())
    assert any(n.endswith("conv0_weight") for n in names) and any(n.endswith("running_var") for n in names)
    with pytest.raises(ValueError):
        vision.get_model("nope")
    for n in ["alexnet", "vgg11_bn", "squeezenet1_1", "densenet121", "mobilenet0_25", "mobilenet_v2_0_25", "inception_v3"]:
        assert len(vision.get_model(n, classes=3).collect_params()) > 0


def test_fused_resnet_trains_on_cpu_reference_path():
    from mxb200.gluon.model_zoo import vision
    net = vision.get_resnet("v1b", 18, classes=4, thumbnail=True, ctx=mx.cpu(), dtype="float32", layout="NHWC")
    tr = gluon.Trainer(net.collect_params(), "nag", {"learning_rate": 0.05, "momentum": 0.9, "wd": 1e-4})
    L = gluon.loss.SoftmaxCrossEntropyLoss()
    x = mx.nd.array(np.random.randn(8, 16, 16, 3))
    y = mx.nd.array(np.random.randint(0, 4, (8,)))
    losses = []
    for _ in range(6):
        with autograd.record():
            l = L(net(x), y)
        l.backward()
        tr.step(8)
        losses.append(float(l.mean().asscalar()))
    assert losses[-1] < losses[0] * 0.5


def test_rnn_layers_and_cells():
    lstm = gluon.rnn.LSTM(6, num_layers=2, bidirectional=True)
    lstm.initialize()
    out = lstm(mx.nd.array(np.random.randn(5, 3, 4)))
    assert out.shape == (5, 3, 12)
    out, st = lstm(mx.nd.array(np.random.randn(5, 3, 4)), lstm.begin_state(3))
    assert st[0].shape == (4, 3, 6) and st[1].shape == (4, 3, 6)
    cell = gluon.rnn.GRUCell(5)
    cell.initialize()
    o, s = cell.unroll(3, mx.nd.array(np.random.randn(2, 3, 4)), layout="NTC", merge_outputs=True)
    assert o.shape == (2, 3, 5)


def test_split_and_load_and_clip_global_norm():
    parts = gluon.utils.split_and_load(mx.nd.array(np.arange(12).reshape(6, 2)), [mx.cpu(0), mx.cpu(1), mx.cpu(2)])
    assert [p.shape for p in parts] == [(2, 2)] * 3 and parts[2].context == mx.cpu(2)
    with pytest.raises(ValueError):
        gluon.utils.split_data(mx.nd.ones((5, 2)), 2)
    arrs = [mx.nd.ones((3,)) * 3, mx.nd.ones((4,)) * 4]
    n = gluon.utils.clip_global_norm(arrs, 1.0)
    assert n == pytest.approx(np.sqrt(27 + 64), rel=1e-5)
    assert np.sqrt(sum((a.asnumpy() ** 2).sum() for a in arrs)) == pytest.approx(1.0, rel=1e-4)


def test_data_loader_and_samplers():
    ds = gluon.data.ArrayDataset(np.arange(20, dtype=np.float32).reshape(10, 2), np.arange(10))
    batches = list(gluon.data.DataLoader(ds, batch_size=4, last_batch="keep"))
    assert [b[0].shape[0] for b in batches] == [4, 4, 2]
    batches = list(gluon.data.DataLoader(ds, batch_size=4, last_batch="discard", num_workers=2))
    assert len(batches) == 2
    s = gluon.data.SplitSampler(10, num_parts=2, part_index=1, shuffle=False)
    assert list(s) == [5, 6, 7, 8, 9]


def test_gluon_bert_model_trains_on_cpu():
    """BERTModel (gluon front-end of models/bert.py): parameters are gluon Parameters, LAMB through gluon.Trainer."""
    import torch
    import mxb200 as mx
    from mxb200 import autograd, gluon
    from mxb200.gluon.model_zoo import bert as zb
    net = zb.BERTModel(vocab=200, hidden=256, layers=1, heads=4, intermediate=512, max_pos=16, dtype="float32", dropout=0.0)
    net.initialize(ctx=mx.cpu())
    assert "word_embed" in "".join(net.collect_params().keys())
    B, S, Mp = 4, 8, 2
    g = torch.Generator().manual_seed(0)
    ids = mx.nd.array(torch.randint(0, 200, (B, S), generator=g).numpy())
    mp = mx.nd.array(torch.randint(0, S, (B, Mp), generator=g).numpy())
    lab = mx.nd.array(torch.randint(0, 200, (B * Mp,), generator=g).numpy())
    tr = gluon.Trainer(net.collect_params(), "lamb", {"learning_rate": 5e-3})
    L = gluon.loss.SoftmaxCrossEntropyLoss()
    losses = []
    for _ in range(6):
        with autograd.record():
            mlm, nsp = net(ids, None, None, mp)
            loss = L(mlm, lab).mean()
        loss.backward()
        tr.step(1)
        losses.append(float(loss.asnumpy().reshape(-1)[0]))
    assert mlm.shape == (B * Mp, 200) and nsp.shape == (B, 8)
    assert losses[-1] < losses[0]
    enc = net(ids)
    assert enc.shape == (B, S, 256)


def test_train_step_eager_path_matches_manual_loop():
    """mx.gluon.TrainStep without capture (CPU) is exactly record / backward / step."""
    import torch
    import mxb200 as mx
    from mxb200 import gluon, autograd

    def build():
        torch.manual_seed(0)
        mx.random.seed(0)
        net = gluon.nn.HybridSequential(prefix="ts_")
        with net.name_scope():
            net.add(gluon.nn.Dense(16, activation="relu", in_units=8), gluon.nn.Dense(4, in_units=16))
        net.initialize(mx.init.Xavier())
        tr = gluon.Trainer(net.collect_params(), "sgd", {"learning_rate": 0.1, "momentum": 0.9})
        return net, tr

    lossf = gluon.loss.SoftmaxCrossEntropyLoss()
    xs = [mx.nd.array(torch.randn(6, 8).numpy()) for _ in range(4)]
    ys = [mx.nd.array(torch.randint(0, 4, (6,)).numpy()) for _ in range(4)]
    net, tr = build()
    step = gluon.TrainStep(net, lossf, tr, batch_size=6)
    a = [float(step(x, y).asnumpy().mean()) for x, y in zip(xs, ys)]
    assert step.replays == 0 and step.eager_calls == 4
    net2, tr2 = build()
    b = []
    for x, y in zip(xs, ys):
        with autograd.record():
            l = lossf(net2(x), y)
        l.backward()
        tr2.step(6)
        b.append(float(l.asnumpy().mean()))
    assert a == b


def test_async_checkpoint_is_written_by_the_engine_and_load_waits(tmp_path):
    import os
    import mxb200 as mx
    from mxb200 import engine
    prefix = str(tmp_path / "ck")
    arg = {"w": mx.nd.ones((64, 64)) * 3, "b": mx.nd.zeros((64,))}
    mx.model.save_checkpoint(prefix, 7, None, arg, {}, async_write=True)
    arg["w"][:] = 0                                        # the snapshot was taken at call time
    _sym, a2, _aux = mx.model.load_checkpoint(prefix, 7)   # waits for the engine's write of exactly this file
    assert float(a2["w"].asnumpy().mean()) == 3.0
    engine.get().wait_all()
    assert os.path.exists(prefix + "-0007.params")
