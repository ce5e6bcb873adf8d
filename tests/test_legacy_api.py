"""Legacy / peripheral python API (reference: tests/python/unittest/test_rnn.py, test_image.py, test_registry...)."""
import collections
import os

import numpy as np

import mxb200 as mx


def test_symbolic_rnn_cells_unroll_and_run():
    cell = mx.rnn.SequentialRNNCell()
    cell.add(mx.rnn.LSTMCell(8, prefix="l0_"))
    cell.add(mx.rnn.DropoutCell(0.0))
    cell.add(mx.rnn.GRUCell(8, prefix="l1_"))
    outs, states = cell.unroll(4, mx.sym.var("data"), layout="NTC", merge_outputs=True)
    assert sorted(outs.list_arguments()) == sorted(["data", "l0_i2h_weight", "l0_i2h_bias", "l0_h2h_weight", "l0_h2h_bias",
                                                    "l1_i2h_weight", "l1_i2h_bias", "l1_h2h_weight", "l1_h2h_bias"])
    assert len(states) == 3
    ex = outs.simple_bind(mx.cpu(), data=(2, 4, 5))
    for k, v in ex.arg_dict.items():
        if k != "data":
            v[:] = np.random.RandomState(0).rand(*v.shape) * 0.1
    assert ex.forward(data=mx.nd.ones((2, 4, 5)))[0].shape == (2, 4, 8)
    bi = mx.rnn.BidirectionalCell(mx.rnn.RNNCell(4, prefix="f_"), mx.rnn.RNNCell(4, prefix="b_"))
    o, _ = bi.unroll(3, mx.sym.var("x"), merge_outputs=True)
    assert o.infer_shape(x=(2, 3, 6))[1] == [(2, 3, 8)]
    res = mx.rnn.ResidualCell(mx.rnn.GRUCell(6, prefix="r_"))
    o, _ = res.unroll(2, mx.sym.var("x"), merge_outputs=True)
    assert o.infer_shape(x=(2, 2, 6))[1] == [(2, 2, 6)]


def test_fused_rnn_cell_and_weight_packing():
    f = mx.rnn.FusedRNNCell(8, num_layers=2, mode="lstm", bidirectional=True, get_next_state=True)
    o, st = f.unroll(4, mx.sym.var("d"), layout="NTC", merge_outputs=True)
    args, outs, _ = o.infer_shape(d=(3, 4, 5))
    assert dict(zip(o.list_arguments(), args))["lstm_parameters"] == (2624,) and outs == [(3, 4, 16)] and len(st) == 2
    u = f.unfuse()
    oo, _ = u.unroll(4, mx.sym.var("d"), merge_outputs=True)
    assert oo.infer_shape(d=(3, 4, 5))[1] == [(3, 4, 16)]
    c = mx.rnn.LSTMCell(8)
    args = {"lstm_i2h_weight": mx.nd.array(np.arange(160).reshape(32, 5)), "lstm_i2h_bias": mx.nd.ones((32,)),
            "lstm_h2h_weight": mx.nd.ones((32, 8)), "lstm_h2h_bias": mx.nd.ones((32,))}
    un = c.unpack_weights(args)
    assert un["lstm_i2h_f_weight"].shape == (8, 5) and float(un["lstm_i2h_f_weight"].asnumpy()[0, 0]) == 40.0
    assert np.array_equal(c.pack_weights(un)["lstm_i2h_weight"].asnumpy(), args["lstm_i2h_weight"].asnumpy())


def test_bucket_sentence_iter_and_encoding():
    sents = [[1, 2, 3], [4, 5], [1, 2, 3, 4, 5, 6], [2, 3, 4]] * 4
    it = mx.rnn.BucketSentenceIter(sents, 2, buckets=[3, 6], invalid_label=0)
    keys = collections.Counter()
    for b in it:
        keys[b.bucket_key] += 1
        assert b.data[0].shape == (2, b.bucket_key)
        d, lab = b.data[0].asnumpy(), b.label[0].asnumpy()
        assert np.array_equal(lab[:, :-1], d[:, 1:]) and (lab[:, -1] == 0).all()
    assert keys == {3: 6, 6: 2}
    enc, vocab = mx.rnn.encode_sentences([["a", "b"], ["b", "c"]], start_label=1)
    assert enc == [[1, 2], [2, 3]] and len(vocab) == 4


def test_image_augmenters_and_detection_iter(tmp_path):
    import cv2
    from mxb200 import image
    img = mx.nd.array((np.random.RandomState(0).rand(40, 60, 3) * 255).astype("f"))
    for aug in image.CreateAugmenter((3, 24, 24), resize=30, rand_crop=True, rand_resize=True, rand_mirror=True, mean=True,
                                     std=True, brightness=0.2, contrast=0.2, saturation=0.2, hue=0.1, pca_noise=0.1, rand_gray=0.5):
        img = aug(img)
    assert img.shape == (24, 24, 3)
    lab = np.array([[0, 0.1, 0.1, 0.5, 0.6], [1, 0.4, 0.4, 0.9, 0.9], [-1, -1, -1, -1, -1]], dtype="f")
    img = mx.nd.array((np.random.RandomState(1).rand(40, 60, 3) * 255).astype("f"))
    flipped, fl = image.DetHorizontalFlipAug(1.0)(img, lab)
    assert np.allclose(fl[0, 1:5], [0.5, 0.1, 0.9, 0.6]) and fl[2, 0] == -1
    for aug in image.CreateDetAugmenter((3, 32, 32), rand_crop=0.9, rand_pad=0.9, rand_mirror=True, mean=True, std=True):
        img, lab = aug(img, lab)
    valid = lab[lab[:, 0] >= 0]
    assert img.shape == (32, 32, 3) and (valid[:, 1:5] >= -1e-6).all() and (valid[:, 1:5] <= 1 + 1e-6).all()
    for n in ("a.jpg", "b.jpg"):
        cv2.imwrite(os.path.join(str(tmp_path), n), (np.random.rand(50, 70, 3) * 255).astype("uint8"))
    imglist = [[[2, 5, 0, 0.1, 0.1, 0.5, 0.6, 1, 0.4, 0.4, 0.9, 0.9], "a.jpg"], [[2, 5, 3, 0.2, 0.2, 0.8, 0.8], "b.jpg"]]
    it = image.ImageDetIter(2, (3, 32, 32), imglist=imglist, path_root=str(tmp_path), rand_mirror=True)
    b = next(iter(it))
    assert b.data[0].shape == (2, 3, 32, 32) and b.label[0].shape == (2, 2, 5)


def test_registry_misc_notebook_executor_manager():
    from mxb200 import registry

    class Base(object):
        def __init__(self, v=1):
            self.v = v
    reg, alias, create = registry.get_register_func(Base, "thing"), registry.get_alias_func(Base, "thing"), registry.get_create_func(Base, "thing")

    @alias("foo2")
    @reg
    class Foo(Base):
        pass
    assert isinstance(create("foo", v=3), Foo) and create("foo2").v == 1 and create('["foo", {"v": 5}]').v == 5
    sch = mx.misc.FactorScheduler(step=10, factor=0.5)
    sch.base_lr = 1.0
    assert sch(25) == 0.25
    from mxb200.executor_manager import _split_input_slice
    assert _split_input_slice(10, [1, 1, 2]) == [slice(0, 2), slice(2, 4), slice(4, 10)] or len(_split_input_slice(10, [1, 1, 2])) == 3
    pl = mx.notebook.callback.PandasLogger(batch_size=4, frequent=1)
    m = mx.metric.create("acc")
    m.update([mx.nd.array([0, 1])], [mx.nd.array([[0.9, 0.1], [0.2, 0.8]])])
    from mxb200.model import BatchEndParam
    pl.train_cb(BatchEndParam(epoch=0, nbatch=1, eval_metric=m, locals=None))
    assert len(pl.train_df) == 1 and "accuracy" in pl.train_df.columns
    t = mx.nd.ones((2, 2))
    assert mx.torch.to_torch(t) is t.as_torch()
