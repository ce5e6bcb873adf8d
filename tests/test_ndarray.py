"""NDArray basics and the .params format (reference tests: tests/python/unittest/test_ndarray.py)."""
import io
import os
import struct

import numpy as np
import pytest
import torch

import mxb200 as mx
from mxb200.ndarray import serialization as ser


def test_arithmetic_and_indexing():
    a = mx.nd.array([[1, 2], [3, 4]])
    b = mx.nd.ones((2, 2))
    np.testing.assert_allclose((a + b * 2 - 1).asnumpy(), [[2, 3], [4, 5]])
    np.testing.assert_allclose((2 / a).asnumpy(), [[2, 1], [2 / 3, 0.5]], rtol=1e-6)
    a[:] = 7
    assert (a.asnumpy() == 7).all()
    a[0, 1] = 3
    assert a[0, 1].asscalar() == 3
    c = mx.nd.arange(0, 10).reshape((2, 5))
    assert c[1].shape == (5,) and c.T.shape == (5, 2)
    assert mx.nd.dot(a, b).shape == (2, 2)
    assert c.reshape((-1,)).shape == (10,) and c.reshape((0, -1)).shape == (2, 5)
    assert mx.nd.zeros((3, 4)).reshape((-3,)).shape == (12,)


def test_copy_context_and_dtype():
    a = mx.nd.array(np.arange(6).reshape(2, 3), dtype="float16")
    assert a.dtype == np.float16 and a.context == mx.cpu(0)
    b = a.as_in_context(mx.cpu(1))
    assert b.context == mx.cpu(1)
    c = mx.nd.zeros((2, 3))
    a.astype("float32").copyto(c)
    np.testing.assert_allclose(c.asnumpy(), np.arange(6).reshape(2, 3))
    bf = mx.nd.ones((2,), dtype="bfloat16")
    assert bf.dtype == mx.bfloat16 and bf.asnumpy().dtype == np.float32


def test_autograd_write_and_add():
    x = mx.nd.array([1.0, 2.0, 3.0])
    x.attach_grad()
    for _ in range(2):
        with mx.autograd.record():
            y = (x * x).sum()
        y.backward()
    np.testing.assert_allclose(x.grad.asnumpy(), [2, 4, 6])      # 'write' overwrites
    z = mx.nd.array([1.0, 2.0])
    z.attach_grad("add")
    for _ in range(2):
        with mx.autograd.record():
            (z * 3).sum().backward()
    np.testing.assert_allclose(z.grad.asnumpy(), [6, 6])          # 'add' accumulates
    with mx.autograd.record():
        w = x * 2
    assert mx.autograd.is_recording() is False and w._t.requires_grad


def test_save_load_roundtrip(tmp_path):
    d = {"arg:w": mx.nd.array(np.random.randn(3, 4)), "aux:m": mx.nd.array(np.arange(5), dtype="int32"),
         "h": mx.nd.array(np.random.randn(2, 2), dtype="float16"), "b": mx.nd.ones((3,), dtype="bfloat16")}
    f = str(tmp_path / "x.params")
    mx.nd.save(f, d)
    back = mx.nd.load(f)
    assert set(back) == set(d)
    for k in d:
        assert back[k].dtype == d[k].dtype and back[k].shape == d[k].shape
        np.testing.assert_array_equal(back[k].asnumpy(), d[k].asnumpy())
    lst = [mx.nd.ones((2,)), mx.nd.zeros((1, 2))]
    mx.nd.save(f, lst)
    back = mx.nd.load(f)
    assert isinstance(back, list) and back[1].shape == (1, 2)


def test_native_writer_matches_python_writer(tmp_path):
    d = {"arg:fc_weight": mx.nd.array(np.random.randn(7, 5)), "aux:bn_moving_var": mx.nd.array(np.random.rand(5))}
    f = str(tmp_path / "n.params")
    mx.nd.save(f, d)          # native path when the host library is built
    with open(f, "rb") as fh:
        native = fh.read()
    assert native == ser.dumps(d)


def test_file_layout_is_reference_format(tmp_path):
    a = mx.nd.array(np.arange(6, dtype=np.float32).reshape(2, 3))
    raw = ser.dumps({"k": a})
    magic, reserved, n = struct.unpack("<QQQ", raw[:24])
    assert (magic, reserved, n) == (0x112, 0, 1)
    nd_magic, stype, ndim = struct.unpack("<IiI", raw[24:36])
    assert nd_magic == 0xF993FAC9 and stype == 0 and ndim == 2
    assert struct.unpack("<qq", raw[36:52]) == (2, 3)
    dev_type, dev_id, flag = struct.unpack("<iii", raw[52:64])
    assert (dev_type, dev_id, flag) == (1, 0, 0)
    assert np.frombuffer(raw[64:88], dtype=np.float32).tolist() == list(range(6))


def test_load_v1_and_legacy_blobs():
    data = np.arange(4, dtype=np.float32)
    v1 = struct.pack("<QQQ", 0x112, 0, 1) + struct.pack("<I", 0xF993FAC8) + struct.pack("<Iqq", 2, 2, 2) + \
        struct.pack("<iii", 1, 0, 0) + data.tobytes() + struct.pack("<Q", 0)
    out = ser.loads(v1)
    assert out[0].shape == (2, 2)
    legacy = struct.pack("<QQQ", 0x112, 0, 1) + struct.pack("<III", 2, 2, 2) + struct.pack("<iii", 1, 0, 0) + \
        data.tobytes() + struct.pack("<Q", 0)
    out = ser.loads(legacy)
    np.testing.assert_array_equal(out[0].asnumpy(), data.reshape(2, 2))
    with pytest.raises(mx.MXNetError):
        ser.loads(b"\x00" * 24)


def test_sparse_roundtrip_and_ops(tmp_path):
    dense = np.zeros((6, 3), np.float32)
    dense[1] = 1
    dense[4] = 2
    rsp = mx.nd.array(dense).tostype("row_sparse")
    assert rsp.indices.asnumpy().tolist() == [1, 4]
    csr = mx.nd.array(dense).tostype("csr")
    f = str(tmp_path / "s.params")
    mx.nd.save(f, {"r": rsp, "c": csr})
    back = mx.nd.load(f)
    np.testing.assert_array_equal(back["r"].asnumpy(), dense)
    np.testing.assert_array_equal(back["c"].asnumpy(), dense)
    out = mx.nd.dot(csr, mx.nd.ones((3, 2)))
    np.testing.assert_allclose(out.asnumpy(), dense @ np.ones((3, 2)))
    kept = mx.nd.sparse.retain(rsp, mx.nd.array([4, 5]))
    assert kept.indices.asnumpy().tolist() == [4]


def test_op_namespace():
    x = mx.nd.array(np.random.randn(4, 10))
    np.testing.assert_allclose(mx.nd.softmax(x).asnumpy().sum(1), np.ones(4), rtol=1e-5)
    assert mx.nd.topk(x, k=3).shape == (4, 3)
    assert mx.nd.one_hot(mx.nd.array([1, 3]), 5).asnumpy()[1, 3] == 1
    assert mx.nd.concat(x, x, dim=1).shape == (4, 20)
    assert len(mx.nd.split(x, num_outputs=2, axis=1)) == 2
    assert mx.nd.pick(x, mx.nd.array([0, 1, 2, 3])).shape == (4,)
    w, b = mx.nd.ones((3, 10)), mx.nd.zeros((3,))
    assert mx.nd.FullyConnected(x, w, b, num_hidden=3).shape == (4, 3)
    img = mx.nd.array(np.random.randn(2, 3, 8, 8))
    k = mx.nd.array(np.random.randn(4, 3, 3, 3))
    assert mx.nd.Convolution(img, k, kernel=(3, 3), pad=(1, 1), num_filter=4, no_bias=True).shape == (2, 4, 8, 8)
    assert mx.nd.Pooling(img, kernel=(2, 2), stride=(2, 2), pool_type="max").shape == (2, 3, 4, 4)
    assert mx.nd.contrib.div_sqrt_dim(x).shape == x.shape


def test_asnumpy_returns_a_private_copy():
    """asnumpy() copies (MXNDArraySyncCopyToCPU); later in-place updates of the NDArray must not show through."""
    a = mx.nd.ones((3,))
    n = a.asnumpy()
    a[:] = 5
    assert n.tolist() == [1, 1, 1] and a.asnumpy().tolist() == [5, 5, 5]


def test_iteration_over_first_axis_terminates():
    """``for row in arr`` uses the __getitem__/IndexError protocol (or __iter__): it must stop after shape[0] rows."""
    a = mx.nd.arange(0, 6).reshape((3, 2))
    assert [r.asnumpy().tolist() for r in a] == [[0.0, 1.0], [2.0, 3.0], [4.0, 5.0]]
    with pytest.raises(IndexError):
        a[5]
