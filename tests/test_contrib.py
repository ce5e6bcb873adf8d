"""contrib / gluon.contrib packages (reference: tests/python/unittest/test_contrib_text.py, test_contrib_svrg_module.py,
test_gluon_contrib.py)."""
import collections
import os

import numpy as np

import mxb200 as mx
from mxb200 import autograd, gluon
from mxb200.contrib import quantization, svrg_optimization, text
from mxb200.gluon.contrib import data as cdata, nn as cnn, rnn as crnn


def test_text_vocab_and_embeddings(tmp_path):
    c = text.utils.count_tokens_from_str("a b b c c c\nd d d d")
    assert c == collections.Counter({"d": 4, "c": 3, "b": 2, "a": 1})
    v = text.Vocabulary(c, most_freq_count=3, reserved_tokens=["<pad>"])
    assert v.idx_to_token == ["<unk>", "<pad>", "d", "c", "b"]
    assert v.to_indices(["c", "zzz"]) == [3, 0] and v.to_tokens([2, 0]) == ["d", "<unk>"]
    p = os.path.join(str(tmp_path), "e.txt")
    open(p, "w").write("a 1 2\nb 3 4\nc 5 6\n")
    e = text.embedding.CustomEmbedding(p)
    assert e.vec_len == 2 and np.allclose(e.get_vecs_by_tokens(["b", "q"]).asnumpy(), [[3, 4], [0, 0]])
    ce = text.embedding.CompositeEmbedding(v, [e, e])
    assert ce.idx_to_vec.shape == (5, 4) and np.allclose(ce.get_vecs_by_tokens("c").asnumpy(), [5, 6, 5, 6])
    e.update_token_vectors("a", mx.nd.array([9.0, 9.0]))
    assert np.allclose(e.get_vecs_by_tokens("a").asnumpy(), [9, 9])


def test_svrg_module_reduces_loss():
    rs = np.random.RandomState(0)
    x = rs.rand(64, 4).astype("f")
    y = x @ np.array([1.0, -2.0, 3.0, 0.5], "f")
    it = mx.io.NDArrayIter(x, y, batch_size=16, label_name="lin_reg_label")
    net = mx.sym.LinearRegressionOutput(mx.sym.FullyConnected(mx.sym.var("data"), num_hidden=1, name="fc"),
                                        mx.sym.var("lin_reg_label"), name="lin_reg")
    m = svrg_optimization.SVRGModule(net, data_names=("data",), label_names=("lin_reg_label",), update_freq=2)
    m.bind(data_shapes=it.provide_data, label_shapes=it.provide_label)
    m.init_params(initializer=mx.init.Uniform(0.01))
    before = dict(m.score(it, "mse"))["mse"]
    m.fit(it, num_epoch=8, eval_metric="mse", optimizer="sgd", optimizer_params=(("learning_rate", 0.2),))
    after = dict(m.score(it, "mse"))["mse"]
    assert after < 0.5 * before


def test_gluon_contrib_layers_and_cells():
    net = cnn.HybridConcurrent(axis=1)
    net.add(gluon.nn.Dense(3), cnn.Identity(), gluon.nn.Dense(2))
    net.initialize()
    assert net(mx.nd.ones((2, 4))).shape == (2, 9)
    cell = crnn.LSTMPCell(8, 3)
    cell.initialize()
    out, st = cell.unroll(5, mx.nd.ones((2, 5, 4)), merge_outputs=True)
    assert out.shape == (2, 5, 3) and [s.shape for s in st] == [(2, 3), (2, 8)]
    cc = crnn.Conv2DLSTMCell((3, 8, 8), 4, (3, 3), (3, 3), i2h_pad=(1, 1))
    cc.initialize()
    o, s = cc(mx.nd.ones((2, 3, 8, 8)), cc.begin_state(batch_size=2))
    assert o.shape == (2, 4, 8, 8) and len(s) == 2
    gg = crnn.Conv1DGRUCell((2, 10), 5, 3, 3)
    gg.initialize()
    assert gg(mx.nd.ones((1, 2, 10)), gg.begin_state(batch_size=1))[0].shape == (1, 5, 8)
    vd = crnn.VariationalDropoutCell(gluon.rnn.LSTMCell(6), drop_inputs=0.5)
    vd.base_cell.initialize()
    with autograd.record():
        o, _ = vd.unroll(3, mx.nd.ones((2, 3, 4)), merge_outputs=True)
    assert o.shape == (2, 3, 6) and "in" in vd._masks       # one mask, reused over the 3 steps
    assert list(cdata.IntervalSampler(7, 3)) == [0, 3, 6, 1, 4, 2, 5]
    assert list(cdata.IntervalSampler(7, 3, rollover=False)) == [0, 3, 6]
    sb = cnn.SyncBatchNorm(in_channels=4)
    sb.initialize()
    with autograd.record():
        y = sb(mx.nd.array(np.random.RandomState(0).rand(2, 4, 3, 3).astype("f")))
    assert abs(float(y.asnumpy().mean())) < 1e-4


def test_quantization_thresholds_and_params():
    a = np.random.RandomState(0).randn(20000).astype("f")
    mn, mx_, div, th = quantization._get_optimal_threshold(a)
    assert 2.0 < th <= max(abs(mn), abs(mx_)) + 1e-6
    q = quantization.quantize_params({"fc_weight": mx.nd.array([[1.0, -2.0], [0.5, 0.25]]), "bn_gamma": mx.nd.ones((2,))})
    assert set(q) == {"fc_weight_quantize", "fc_weight_min", "fc_weight_max", "bn_gamma"}
    assert q["fc_weight_quantize"].asnumpy().tolist() == [[64, -127], [32, 16]]
