"""Data iterators / RecordIO (reference tests: tests/python/unittest/test_io.py, test_recordio.py)."""
import numpy as np
import pytest

import mxb200 as mx


def test_ndarray_iter_pad_discard_rollover():
    data = np.arange(100).reshape(25, 4).astype(np.float32)
    label = np.arange(25).astype(np.float32)
    it = mx.io.NDArrayIter(data, label, 10, last_batch_handle="pad")
    batches = list(it)
    assert len(batches) == 3 and batches[-1].pad == 5
    np.testing.assert_array_equal(batches[-1].data[0].asnumpy()[5:], data[:5])     # padded by wrapping around
    assert it.provide_data[0].shape == (10, 4) and it.provide_label[0].name == "softmax_label"
    it = mx.io.NDArrayIter(data, label, 10, last_batch_handle="discard")
    assert len(list(it)) == 2
    it = mx.io.NDArrayIter({"a": data, "b": data * 2}, label, 5)
    b = next(iter(it))
    assert [d.name for d in it.provide_data] == ["a", "b"] and len(b.data) == 2


def test_resize_and_prefetching_iter():
    data = np.random.randn(20, 3).astype(np.float32)
    base = mx.io.NDArrayIter(data, np.zeros(20, np.float32), 5)
    r = mx.io.ResizeIter(base, 7)
    assert sum(1 for _ in r) == 7           # wraps around the underlying 4 batches
    r.reset()
    assert sum(1 for _ in r) == 7
    p = mx.io.PrefetchingIter(mx.io.NDArrayIter(data, np.zeros(20, np.float32), 5))
    n = 0
    for batch in p:
        assert batch.data[0].shape == (5, 3)
        n += 1
    assert n == 4
    p.reset()
    assert sum(1 for _ in p) == 4


def test_csv_iter_sharding(tmp_path):
    d = np.arange(60).reshape(20, 3).astype(np.float32)
    f = str(tmp_path / "d.csv")
    np.savetxt(f, d, delimiter=",")
    it0 = mx.io.CSVIter(data_csv=f, data_shape=(3,), batch_size=5, num_parts=2, part_index=0)
    it1 = mx.io.CSVIter(data_csv=f, data_shape=(3,), batch_size=5, num_parts=2, part_index=1)
    a = np.concatenate([b.data[0].asnumpy() for b in it0])
    b = np.concatenate([b.data[0].asnumpy() for b in it1])
    np.testing.assert_array_equal(a, d[:10])
    np.testing.assert_array_equal(b, d[10:])


def test_recordio_roundtrip_with_embedded_magic(tmp_path):
    f = str(tmp_path / "a.rec")
    w = mx.recordio.MXRecordIO(f, "w")
    payloads = [b"hello", b"x" * 1001, b"ab" + b"\x0a\x23\xd7\xce" * 3 + b"tail", b""]
    for p in payloads:
        w.write(p)
    w.close()
    r = mx.recordio.MXRecordIO(f, "r")
    got = []
    while True:
        s = r.read()
        if s is None:
            break
        got.append(s)
    assert got[:3] == payloads[:3]


def test_indexed_recordio_and_pack(tmp_path):
    rec, idx = str(tmp_path / "i.rec"), str(tmp_path / "i.idx")
    w = mx.recordio.MXIndexedRecordIO(idx, rec, "w")
    for i in range(5):
        w.write_idx(i, mx.recordio.pack(mx.recordio.IRHeader(0, float(i), i, 0), b"payload%d" % i))
    w.close()
    r = mx.recordio.MXIndexedRecordIO(idx, rec, "r")
    assert r.keys == [0, 1, 2, 3, 4]
    h, s = mx.recordio.unpack(r.read_idx(3))
    assert h.label == 3.0 and s == b"payload3"
    h, s = mx.recordio.unpack(mx.recordio.pack(mx.recordio.IRHeader(0, [1.0, 2.0, 3.0], 7, 0), b"zz"))
    assert h.flag == 3 and h.label.tolist() == [1.0, 2.0, 3.0] and s == b"zz"


def test_image_record_iter_pipeline(tmp_path):
    rec, idx = str(tmp_path / "img.rec"), str(tmp_path / "img.idx")
    w = mx.recordio.MXIndexedRecordIO(idx, rec, "w")
    rng = np.random.RandomState(0)
    for i in range(10):
        img = rng.randint(0, 255, (40, 48, 3), dtype=np.uint8)
        w.write_idx(i, mx.recordio.pack_img(mx.recordio.IRHeader(0, float(i % 3), i, 0), img, quality=95))
    w.close()
    it = mx.io.ImageRecordIter(path_imgrec=rec, path_imgidx=idx, data_shape=(3, 32, 32), batch_size=4, rand_crop=True,
                               rand_mirror=True, preprocess_threads=2)
    shapes = [(b.data[0].shape, b.pad) for b in it]
    assert shapes == [((4, 3, 32, 32), 0), ((4, 3, 32, 32), 0), ((4, 3, 32, 32), 2)]
    it.reset()
    assert sum(1 for _ in it) == 3
    half = mx.io.ImageRecordIter(path_imgrec=rec, path_imgidx=idx, data_shape=(3, 32, 32), batch_size=5, num_parts=2,
                                 part_index=1)
    labs = np.concatenate([b.label[0].asnumpy() for b in half])
    np.testing.assert_array_equal(labs, [float(i % 3) for i in range(5, 10)])


def test_metrics():
    acc = mx.metric.create("acc")
    acc.update([mx.nd.array([0, 1, 1])], [mx.nd.array([[0.9, 0.1], [0.2, 0.8], [0.7, 0.3]])])
    assert acc.get() == ("accuracy", pytest.approx(2 / 3))
    topk = mx.metric.TopKAccuracy(top_k=2)
    topk.update([mx.nd.array([2, 0])], [mx.nd.array([[0.1, 0.5, 0.4], [0.1, 0.5, 0.4]])])
    assert topk.get()[1] == 0.5
    ce = mx.metric.CrossEntropy()
    ce.update([mx.nd.array([0, 1])], [mx.nd.array([[0.5, 0.5], [0.25, 0.75]])])
    assert ce.get()[1] == pytest.approx(-(np.log(0.5) + np.log(0.75)) / 2, rel=1e-5)
    comp = mx.metric.create(["acc", "mse"])
    comp.update([mx.nd.array([1.0, 0.0])], [mx.nd.array([1.0, 1.0])])
    names, vals = comp.get()
    assert names == ["accuracy", "mse"] and vals[1] == pytest.approx(0.5)
    ppl = mx.metric.Perplexity(ignore_label=None)
    ppl.update([mx.nd.array([0, 1])], [mx.nd.array([[0.5, 0.5], [0.5, 0.5]])])
    assert ppl.get()[1] == pytest.approx(2.0)
    f1 = mx.metric.F1()
    f1.update([mx.nd.array([1, 0, 1, 1])], [mx.nd.array([[0.1, 0.9], [0.8, 0.2], [0.6, 0.4], [0.3, 0.7]])])
    assert f1.get()[1] == pytest.approx(0.8)
    custom = mx.metric.np(lambda l, p: float(np.abs(l - p).sum()))
    custom.update([mx.nd.array([1.0])], [mx.nd.array([3.0])])
    assert custom.get()[1] == 2.0


def test_device_prefetcher_cpu_path():
    import torch
    import mxb200 as mx
    imgs = [torch.randint(0, 256, (2, 4, 4, 3), dtype=torch.uint8) for _ in range(3)]
    labs = [torch.tensor([i, i + 1]) for i in range(3)]
    pf = mx.io.DevicePrefetcher(list(zip(imgs, labs)), mx.cpu(), mean=(1.0, 2.0, 3.0), std=(2.0, 2.0, 2.0), dtype="float32")
    got = list(pf)
    assert len(got) == 3
    for (x, y), im, lb in zip(got, imgs, labs):
        ref = (im.float() - torch.tensor([1.0, 2.0, 3.0])) / 2.0
        assert torch.allclose(x.as_torch(), ref) and torch.equal(y.as_torch(), lb)
