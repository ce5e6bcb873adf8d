"""Symbol / Executor / Module (reference tests: test_symbol.py, test_executor.py, test_module.py, train/test_mlp.py)."""
import json
import os

import numpy as np
import pytest

import mxb200 as mx


def _mlp():
    data = mx.sym.var("data")
    fc1 = mx.sym.FullyConnected(data, name="fc1", num_hidden=32)
    act = mx.sym.Activation(fc1, name="relu1", act_type="relu")
    fc2 = mx.sym.FullyConnected(act, name="fc2", num_hidden=4)
    return mx.sym.SoftmaxOutput(fc2, name="softmax")


def _toy(n=256, seed=0):
    rng = np.random.RandomState(seed)
    X = rng.randn(n, 10).astype(np.float32)
    Y = (X @ rng.randn(10, 4)).argmax(1).astype(np.float32)
    return X, Y


def test_symbol_compose_infer_json(tmp_path):
    net = _mlp()
    assert net.list_arguments() == ["data", "fc1_weight", "fc1_bias", "fc2_weight", "fc2_bias", "softmax_label"]
    assert net.list_outputs() == ["softmax_output"]
    arg, out, aux = net.infer_shape(data=(8, 10))
    assert arg == [(8, 10), (32, 10), (32,), (4, 32), (4,), (8,)] and out == [(8, 4)] and aux == []
    f = str(tmp_path / "net-symbol.json")
    net.save(f)
    g = json.load(open(f))
    assert {"nodes", "arg_nodes", "heads", "node_row_ptr"} <= set(g)
    net2 = mx.sym.load(f)
    assert net2.list_arguments() == net.list_arguments()
    assert net2.infer_shape(data=(8, 10))[1] == [(8, 4)]
    internals = net.get_internals()
    assert "fc1_output" in internals.list_outputs()
    conv = mx.sym.Convolution(mx.sym.var("x"), kernel=(3, 3), num_filter=8, pad=(1, 1), name="c")
    bn = mx.sym.BatchNorm(conv, name="bn")
    assert bn.list_auxiliary_states() == ["bn_moving_mean", "bn_moving_var"]
    a, o, x = bn.infer_shape(x=(2, 3, 5, 5))
    assert o == [(2, 8, 5, 5)] and x == [(8,), (8,)]
    s = (mx.sym.var("a") + mx.sym.var("b") * 2.0 - 1.0)
    r = s.eval_nd({"a": mx.nd.ones((2,)), "b": mx.nd.ones((2,))})[0]
    np.testing.assert_allclose(r.asnumpy(), [2.0, 2.0])


def test_executor_forward_backward_matches_numpy():
    x, w = mx.sym.var("x"), mx.sym.var("w")
    y = mx.sym.FullyConnected(x, weight=w, no_bias=True, num_hidden=3, name="fc")
    loss = mx.sym.MakeLoss(mx.sym.sum(mx.sym.square(y)))
    xv, wv = np.random.randn(4, 5).astype(np.float32), np.random.randn(3, 5).astype(np.float32)
    ex = loss.bind(mx.cpu(), {"x": mx.nd.array(xv), "w": mx.nd.array(wv)}, grad_req={"x": "null", "w": "write"})
    out = ex.forward(is_train=True)[0]
    ex.backward()
    np.testing.assert_allclose(out.asnumpy(), ((xv @ wv.T) ** 2).sum(), rtol=1e-4)
    np.testing.assert_allclose(ex.grad_dict["w"].asnumpy(), 2 * (xv @ wv.T).T @ xv, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("ctxs", [[mx.cpu(0)], [mx.cpu(0), mx.cpu(1)]])
@pytest.mark.parametrize("kv", ["local", "device"])
def test_module_fit_converges_and_checkpoints(tmp_path, ctxs, kv):
    X, Y = _toy()
    it = mx.io.NDArrayIter(X, Y, batch_size=32, shuffle=True)
    mod = mx.mod.Module(_mlp(), context=ctxs)
    prefix = str(tmp_path / "mlp")
    mod.fit(it, num_epoch=8, optimizer="sgd", optimizer_params={"learning_rate": 0.3, "momentum": 0.9}, kvstore=kv,
            initializer=mx.init.Xavier(), epoch_end_callback=mx.callback.do_checkpoint(prefix),
            batch_end_callback=mx.callback.Speedometer(32, 2))
    acc = mod.score(it, "acc")[0][1]
    assert acc > 0.85
    assert os.path.exists(prefix + "-symbol.json") and os.path.exists(prefix + "-0008.params")
    sym, arg, aux = mx.model.load_checkpoint(prefix, 8)
    assert set(arg) == {"fc1_weight", "fc1_bias", "fc2_weight", "fc2_bias"}
    m2 = mx.mod.Module.load(prefix, 8)
    m2.bind(it.provide_data, it.provide_label, for_training=False)
    assert m2.score(it, "acc")[0][1] == pytest.approx(acc)
    pred = m2.predict(it)
    assert pred.shape == (256, 4)


def test_module_multi_device_equals_single_device():
    X, Y = _toy(64)
    outs = []
    for ctxs in ([mx.cpu(0)], [mx.cpu(0), mx.cpu(1), mx.cpu(2), mx.cpu(3)]):
        mx.random.seed(5)
        np.random.seed(5)
        it = mx.io.NDArrayIter(X, Y, batch_size=32)
        mod = mx.mod.Module(_mlp(), context=ctxs)
        mod.bind(it.provide_data, it.provide_label)
        mod.init_params(mx.init.Constant(0.05))
        mod.init_optimizer(kvstore="device", optimizer="sgd", optimizer_params={"learning_rate": 0.1})
        for batch in it:
            mod.forward_backward(batch)
            mod.update()
        outs.append(mod.get_params()[0]["fc2_weight"].asnumpy())
    np.testing.assert_allclose(outs[0], outs[1], rtol=1e-4, atol=1e-5)


def test_module_save_load_optimizer_states_and_reshape(tmp_path):
    X, Y = _toy(64)
    it = mx.io.NDArrayIter(X, Y, batch_size=16)
    mod = mx.mod.Module(_mlp())
    mod.bind(it.provide_data, it.provide_label)
    mod.init_params()
    mod.init_optimizer(optimizer="sgd", optimizer_params={"learning_rate": 0.1, "momentum": 0.9})
    b = next(iter(it))
    mod.forward_backward(b)
    mod.update()
    prefix = str(tmp_path / "m")
    mod.save_checkpoint(prefix, 1, save_optimizer_states=True)
    assert os.path.exists(prefix + "-0001.states")
    m2 = mx.mod.Module.load(prefix, 1, load_optimizer_states=True)
    m2.bind(it.provide_data, it.provide_label)
    m2.init_optimizer(optimizer="sgd", optimizer_params={"learning_rate": 0.1, "momentum": 0.9})
    m2.forward_backward(b)
    m2.update()
    big = mx.io.DataBatch([mx.nd.array(np.random.randn(8, 10))], [mx.nd.zeros((8,))])
    mod.forward(big, is_train=False)                  # triggers reshape to a new batch size
    assert mod.get_outputs()[0].shape == (8, 4)


def test_bucketing_and_sequential_modules():
    def sym_gen(key):
        data = mx.sym.var("data")
        fc = mx.sym.FullyConnected(data, name="fc", num_hidden=4, flatten=True)
        return mx.sym.SoftmaxOutput(fc, name="softmax"), ("data",), ("softmax_label",)
    bm = mx.mod.BucketingModule(sym_gen, default_bucket_key=10)
    bm.bind([("data", (8, 10))], [("softmax_label", (8,))])
    bm.init_params()
    bm.init_optimizer()
    batch = mx.io.DataBatch([mx.nd.ones((8, 10))], [mx.nd.zeros((8,))], bucket_key=10,
                            provide_data=[("data", (8, 10))], provide_label=[("softmax_label", (8,))])
    bm.forward_backward(batch)
    bm.update()
    assert bm.get_outputs()[0].shape == (8, 4)
    m1 = mx.mod.Module(mx.sym.FullyConnected(mx.sym.var("data"), num_hidden=6, name="a"), label_names=None)
    m2 = mx.mod.Module(mx.sym.SoftmaxOutput(mx.sym.FullyConnected(mx.sym.var("data"), num_hidden=3, name="b"), name="softmax"))
    seq = mx.mod.SequentialModule().add(m1).add(m2, take_labels=True, auto_wiring=True)
    seq.bind([("data", (4, 5))], [("softmax_label", (4,))])
    seq.init_params()
    seq.init_optimizer()
    seq.forward_backward(mx.io.DataBatch([mx.nd.ones((4, 5))], [mx.nd.zeros((4,))]))
    seq.update()
    assert seq.get_outputs()[0].shape == (4, 3)


def test_hybridblock_export_and_symbolblock_import(tmp_path):
    from mxb200.gluon import nn
    net = nn.HybridSequential(prefix="net_")
    with net.name_scope():
        net.add(nn.Dense(8, activation="relu", in_units=5), nn.Dense(3, in_units=8))
    net.initialize()
    net.hybridize()
    x = mx.nd.array(np.random.randn(2, 5))
    y = net(x)
    prefix = str(tmp_path / "exp")
    net.export(prefix)
    assert os.path.exists(prefix + "-symbol.json") and os.path.exists(prefix + "-0000.params")
    blk = mx.gluon.SymbolBlock.imports(prefix + "-symbol.json", ["data"], prefix + "-0000.params")
    np.testing.assert_allclose(blk(x).asnumpy(), y.asnumpy(), rtol=1e-5, atol=1e-6)


def test_model_parallel_group2ctx_placement():
    """ctx_group attribute + group2ctx at bind (reference: tests/python/unittest/test_model_parallel.py)."""
    import mxb200 as mx
    with mx.AttrScope(ctx_group="dev1"):
        a = mx.sym.var("a")
        b = a * 2
    with mx.AttrScope(ctx_group="dev2"):
        c = b + 1
    ex = c.bind(mx.cpu(), {"a": mx.nd.ones((2,))}, group2ctx={"dev1": mx.cpu(0), "dev2": mx.cpu(1)})
    out = ex.forward()[0]
    assert out.asnumpy().tolist() == [3.0, 3.0] and out.context == mx.cpu(1)
    c2 = mx.sym.load_json(c.tojson())                     # placement attributes survive serialisation
    assert c2.bind(mx.cpu(), {"a": mx.nd.ones((2,))}, group2ctx={"dev2": mx.cpu(3)}).forward()[0].context == mx.cpu(3)


def test_symbol_fusion_plan_groups_and_cpu_fallback(monkeypatch):
    """The executor's graph pass groups [1x1 Convolution ->] BatchNorm [-> add] [-> relu]; when the fused kernels do not
    apply (here: CPU) a group runs op by op and gives bit-identical results to MXB200_GRAPH_FUSION=0."""
    def block(data, name, nf, down):
        c1 = mx.sym.Convolution(data, kernel=(1, 1), num_filter=nf // 4, no_bias=True, name=name + "_c1")
        a1 = mx.sym.Activation(mx.sym.BatchNorm(c1, fix_gamma=False, name=name + "_bn1"), act_type="relu")
        c2 = mx.sym.Convolution(a1, kernel=(3, 3), pad=(1, 1), num_filter=nf // 4, no_bias=True, name=name + "_c2")
        a2 = mx.sym.Activation(mx.sym.BatchNorm(c2, fix_gamma=False, name=name + "_bn2"), act_type="relu")
        c3 = mx.sym.Convolution(a2, kernel=(1, 1), num_filter=nf, no_bias=True, name=name + "_c3")
        b3 = mx.sym.BatchNorm(c3, fix_gamma=False, name=name + "_bn3")
        sc = data
        if down:
            sc = mx.sym.BatchNorm(mx.sym.Convolution(data, kernel=(1, 1), num_filter=nf, no_bias=True, name=name + "_sc"),
                                  fix_gamma=False, name=name + "_scbn")
        return mx.sym.Activation(b3 + sc, act_type="relu")
    net = block(block(mx.sym.var("data"), "u1", 64, True), "u2", 64, False)
    for sym in (net, mx.sym.load_json(net.tojson())):           # attributes are strings after a JSON round trip
        deferred, tails = sym._fusion_plan(sym._topo())
        groups = sorted((g[0].name if g[0] is not None else "-", g[1].name, g[2] is not None, g[4]) for g in tails.values())
        assert groups == [("-", "u1_bn2", False, True), ("-", "u2_bn2", False, True), ("u1_c1", "u1_bn1", False, True),
                          ("u1_c3", "u1_bn3", True, True), ("u1_sc", "u1_scbn", False, False),
                          ("u2_c1", "u2_bn1", False, True), ("u2_c3", "u2_bn3", True, True)]
        assert len(deferred) == 13
    # a BatchNorm whose output is also a graph output is never grouped
    bn = mx.sym.BatchNorm(mx.sym.var("x"), name="b")
    both = mx.sym.Group([mx.sym.Activation(bn, act_type="relu"), bn])
    assert both._fusion_plan(both._topo())[1] == {}
    ex = net.simple_bind(mx.cpu(), data=(2, 16, 8, 8))
    rs = np.random.RandomState(0)
    for k, v in ex.arg_dict.items():
        v[:] = mx.nd.array(rs.randn(*v.shape).astype("f") * 0.1)
    ex.forward(is_train=True)
    ex.backward(mx.nd.ones(ex.outputs[0].shape))
    o1, g1 = ex.outputs[0].asnumpy(), ex.grad_dict["u1_c1_weight"].asnumpy()
    monkeypatch.setenv("MXB200_GRAPH_FUSION", "0")
    ex.forward(is_train=True)
    ex.backward(mx.nd.ones(ex.outputs[0].shape))
    assert np.array_equal(o1, ex.outputs[0].asnumpy()) and np.array_equal(g1, ex.grad_dict["u1_c1_weight"].asnumpy())


def test_fit_epoch_begin_callback_runs_before_each_epoch():
    """fork addition to Module.fit (python/mxnet/module/base_module.py:415,537-539)."""
    x = np.random.rand(32, 6).astype("f")
    y = (x.sum(1) > 3).astype("f")
    it = mx.io.NDArrayIter(x, y, batch_size=8)
    net = mx.sym.SoftmaxOutput(mx.sym.FullyConnected(mx.sym.var("data"), num_hidden=2, name="fc"), name="softmax")
    seen, ends = [], []
    mod = mx.mod.Module(net)
    mod.fit(it, num_epoch=3, epoch_begin_callback=lambda epoch, sym, arg, aux: seen.append((epoch, len(ends))),
            epoch_end_callback=lambda epoch, sym, arg, aux: ends.append(epoch))
    assert seen == [(0, 0), (1, 1), (2, 2)]


def test_module_trains_when_write_hooks_are_taken_over():
    """The fused GPU update path (Module._maybe_fuse) removes the parameters' 'write' hooks because its kernel clears the
    gradient arena itself.  The executor's "no gradient reached this argument -> zero" rule must not mistake that for a
    missing gradient (it once did: every gradient was zeroed after backward and the GPU Module stopped learning)."""
    import numpy as np
    import mxb200 as mx
    rs = np.random.RandomState(0)
    x = rs.rand(256, 1, 12, 12).astype("f")
    y = (x.mean(axis=(1, 2, 3)) > 0.5).astype("f")
    x[y == 1, :, :6, :] += 0.5
    it = mx.io.NDArrayIter(x, y, batch_size=32, shuffle=True)
    d = mx.sym.var("data")
    n = mx.sym.Convolution(d, kernel=(3, 3), num_filter=8, name="c1")
    n = mx.sym.BatchNorm(n, name="bn1")
    n = mx.sym.Activation(n, act_type="relu")
    n = mx.sym.Pooling(n, kernel=(2, 2), stride=(2, 2), pool_type="max")
    n = mx.sym.FullyConnected(mx.sym.flatten(n), num_hidden=2, name="fc")
    net = mx.sym.SoftmaxOutput(n, name="softmax")
    mod = mx.mod.Module(net, context=[mx.cpu(0)])
    mod.bind(it.provide_data, it.provide_label)
    mod.init_params(mx.init.Xavier())
    mod.init_optimizer(optimizer="sgd", optimizer_params={"learning_rate": 0.1, "momentum": 0.9})
    ex = mod._exec_group.execs[0]
    for name in mod._exec_group.param_names:
        a = ex.arg_dict[name]
        if a._grad_hook is not None:
            a._grad_hook.remove()
            a._grad_hook = None
    for _ in range(6):
        it.reset()
        for b in it:
            mod.forward_backward(b)
            assert float(abs(ex.grad_dict["fc_weight"].asnumpy()).sum()) > 0
            mod.update()
            for name in mod._exec_group.param_names:
                ex.grad_dict[name][:] = 0
    assert dict(mod.score(it, "acc"))["accuracy"] > 0.9


def test_executor_write_request_zeroes_unreached_gradients():
    import numpy as np
    import mxb200 as mx
    x = mx.sym.var("x")
    y = mx.sym.shape_array(x)
    g = mx.nd.array(np.full((3, 4), 7.0))
    exe = y.bind(mx.cpu(), args={"x": mx.nd.ones((3, 4))}, args_grad={"x": g})
    exe.forward(is_train=True)
    exe.backward()
    assert float(abs(g.asnumpy()).sum()) == 0.0
    # two variables of the same name are kept apart by position
    a, b = mx.sym.var("data"), mx.sym.var("data")
    s = mx.sym.maximum(a, b) + mx.sym.minimum(a, b)
    ga, gb = mx.nd.zeros((2, 2)), mx.nd.zeros((2, 2))
    exe = s.bind(mx.cpu(), args=[mx.nd.full((2, 2), 2.0), mx.nd.full((2, 2), 3.0)], args_grad=[ga, gb])
    out = exe.forward(is_train=True)[0].asnumpy()
    assert (out == 5.0).all()
    exe.backward(mx.nd.ones((2, 2)))
    assert (ga.asnumpy() == 1.0).all() and (gb.asnumpy() == 1.0).all()


def test_module_sparse_grad_embedding_trains_with_row_sparse_gradients():
    """Embedding(sparse_grad=True): the executor hands the weight gradient out as a row_sparse array (rows = looked-up ids),
    the updater / kvstore apply it lazily; backward() may be called twice on one forward()."""
    import numpy as np
    import mxb200 as mx
    rs = np.random.RandomState(0)
    x = rs.randint(0, 50, size=(256, 4)).astype("f")
    y = (x[:, 0] % 2).astype("f")
    for kv in (None, "local"):
        it = mx.io.NDArrayIter(x, y, batch_size=32, shuffle=True)
        e = mx.sym.Embedding(mx.sym.var("data"), input_dim=50, output_dim=8, sparse_grad=True, name="emb")
        net = mx.sym.SoftmaxOutput(mx.sym.FullyConnected(mx.sym.flatten(e), num_hidden=2, name="fc"), name="softmax")
        mod = mx.mod.Module(net, context=[mx.cpu(0)])
        mod.fit(it, num_epoch=25, optimizer="sgd", optimizer_params={"learning_rate": 0.5, "momentum": 0.9}, kvstore=kv,
                initializer=mx.init.Xavier())
        g = mod._exec_group.execs[0].grad_dict["emb_weight"]
        assert g.stype == "row_sparse" and 0 < g.indices.shape[0] <= 50
        assert dict(mod.score(it, "acc"))["accuracy"] > 0.95
    ex = net.simple_bind(mx.cpu(), data=(8, 4), grad_req={"data": "null", "softmax_label": "null", "emb_weight": "write",
                                                          "fc_weight": "write", "fc_bias": "write"})
    ex.arg_dict["data"][:] = x[:8]
    ex.arg_dict["emb_weight"][:] = mx.nd.array(rs.rand(50, 8))
    ex.forward(is_train=True)
    ex.backward()
    first = ex.grad_dict["emb_weight"].asnumpy().copy()
    ex.backward()
    assert np.allclose(first, ex.grad_dict["emb_weight"].asnumpy())
