"""Operator / executor semantics pinned down while running the reference's own unit tests against this package
(tools/run_reference_tests.py).  Each case states the MXNet 1.3 behaviour it protects; expected values are computed with
numpy / plain torch here, nothing is taken from the reference tests."""
import numpy as np
import pytest
import torch

import mxb200 as mx
from mxb200.base import MXNetError


def test_upsampling_nearest_multi_input_concat_and_sum():
    a = mx.nd.array(np.arange(16, dtype="f").reshape(1, 1, 4, 4))
    b = mx.nd.array(np.arange(4, dtype="f").reshape(1, 1, 2, 2))
    out = mx.nd.UpSampling(a, b, scale=2, sample_type="nearest", num_args=2)
    assert out.shape == (1, 2, 8, 8)
    an, bn = a.asnumpy(), b.asnumpy()
    assert (out.asnumpy()[:, 0] == an[:, 0].repeat(2, 1).repeat(2, 2)).all()
    assert (out.asnumpy()[:, 1] == bn[:, 0].repeat(4, 1).repeat(4, 2)).all()
    s = mx.nd.UpSampling(a, b, scale=2, sample_type="nearest", num_args=2, multi_input_mode="sum")
    assert s.shape == (1, 1, 8, 8)


def test_prelu_gamma_broadcast_rules():
    x = np.random.RandomState(0).uniform(-1, 1, (3, 4, 2, 2)).astype("f")
    per_channel = np.array([0.1, 0.2, 0.3, 0.4], "f")
    y = mx.nd.LeakyReLU(mx.nd.array(x), gamma=mx.nd.array(per_channel), act_type="prelu").asnumpy()
    assert np.allclose(y, np.where(x > 0, x, x * per_channel.reshape(1, 4, 1, 1)))
    full = np.tile(per_channel, (3, 1))                       # (N, C): lines up with the leading dimensions
    y = mx.nd.LeakyReLU(mx.nd.array(x), gamma=mx.nd.array(full), act_type="prelu").asnumpy()
    assert np.allclose(y, np.where(x > 0, x, x * full.reshape(3, 4, 1, 1)))


def test_ctc_label_length_is_first_padding_and_frames_past_length_are_ignored():
    rs = np.random.RandomState(1)
    acts = rs.randn(6, 2, 5).astype("f")
    labels = np.array([[1, 2, 0, 3], [2, 2, 1, 0]], "f")     # blank first: 0 pads, the first 0 ends the sequence
    got = mx.nd.CTCLoss(mx.nd.array(acts), mx.nd.array(labels)).asnumpy()
    lp = torch.log_softmax(torch.tensor(acts), 2)
    ref = torch.nn.functional.ctc_loss(lp, torch.tensor([[1, 2, 0, 0], [2, 2, 1, 0]]), torch.tensor([6, 6]), torch.tensor([2, 3]),
                                       blank=0, reduction="none")
    assert np.allclose(got, ref.numpy(), atol=1e-5)
    acts_nan = acts.copy()
    acts_nan[4:] = np.nan
    x = mx.nd.array(acts_nan)
    x.attach_grad()
    with mx.autograd.record():
        l = mx.nd.CTCLoss(x, mx.nd.array(labels), data_lengths=mx.nd.array([4, 4]), use_data_lengths=True)
    l.backward()
    assert np.isfinite(l.asnumpy()).all() and (x.grad.asnumpy()[4:] == 0).all()


def test_integer_modulo_by_zero_is_zero_and_reverse_reshape_codes():
    a = mx.nd.array([5, -5, 7], dtype="int32")
    b = mx.nd.array([3, 3, 0], dtype="int32")
    assert mx.nd.modulo(a, b).asnumpy().tolist() == [2, 1, 0]
    x = mx.nd.zeros((1, 2, 3, 64))
    assert mx.nd.reshape(x, shape=(-2, -1, 16, -4), reverse=True).shape == (1, 2, 3, 4, 16)
    assert mx.nd.reshape(mx.nd.zeros((2, 3, 5, 5)), shape=(0, -1), reverse=True).shape == (5, 30)
    assert mx.nd.Reshape(mx.nd.zeros((2, 3, 5, 5)), target_shape=(2, 0)).shape == (2, 75)


def test_roi_pooling_and_roi_align_against_direct_evaluation():
    rs = np.random.RandomState(2)
    x = rs.rand(2, 3, 8, 8).astype("f")
    rois = np.array([[0, 1, 1, 6, 6], [1, 0, 2, 3, 7]], "f")
    out = mx.nd.ROIPooling(mx.nd.array(x), mx.nd.array(rois), pooled_size=(2, 2), spatial_scale=1.0).asnumpy()
    assert out.shape == (2, 3, 2, 2)
    assert np.allclose(out[0, :, 0, 0], x[0, :, 1:4, 1:4].max(axis=(1, 2)))       # roi 6x6 from (1,1): bins of 3x3
    assert np.allclose(out[0, :, 1, 1], x[0, :, 4:7, 4:7].max(axis=(1, 2)))
    al = mx.nd.contrib.ROIAlign(mx.nd.array(x), mx.nd.array([[0, 0, 0, 4, 4]]), pooled_size=(2, 2), spatial_scale=1.0,
                                sample_ratio=1).asnumpy()
    # one sample per bin at the bin centre (1, 1) / (1, 3) / ...: an integer grid point, i.e. the pixel itself
    assert np.allclose(al[0, :, 0, 0], x[0, :, 1, 1]) and np.allclose(al[0, :, 1, 0], x[0, :, 3, 1])
    d = mx.nd.array(x)
    r = mx.nd.array([[0, 0.5, 0.5, 5.2, 6.1]])
    d.attach_grad()
    r.attach_grad()
    with mx.autograd.record():
        y = mx.nd.contrib.ROIAlign(d, r, pooled_size=(3, 3), spatial_scale=1.0, sample_ratio=2)
    y.backward()
    assert abs(d.grad.asnumpy().sum() - 3 * 9) < 1e-3 and (r.grad.asnumpy() == 0).all()   # bilinear weights sum to 1


def test_custom_op_with_aux_state_keyword_inputs_and_legacy_numpy_op():
    class Sqr(mx.operator.CustomOp):
        def forward(self, is_train, req, in_data, out_data, aux):
            aux[0][:] = 1
            self.assign(out_data[0], req[0], in_data[0] * in_data[0])

        def backward(self, req, out_grad, in_data, out_data, in_grad, aux):
            assert (aux[0].asnumpy() == 1).all()
            self.assign(in_grad[0], req[0], 2 * in_data[0] * out_grad[0])

    @mx.operator.register("parity_sqr")
    class SqrProp(mx.operator.CustomOpProp):
        def list_auxiliary_states(self):
            return ["aux"]

        def infer_shape(self, in_shape):
            return in_shape, [in_shape[0]], [in_shape[0]]

        def create_operator(self, ctx, shapes, dtypes):
            return Sqr()

    sym = mx.sym.Custom(data=mx.sym.var("data"), name="sq", op_type="parity_sqr")
    assert sym.list_arguments() == ["data"] and sym.list_auxiliary_states() == ["sq_aux"]
    assert sym.infer_shape(data=(4, 5))[2] == [(4, 5)]
    x = mx.nd.array(np.random.RandomState(3).rand(4, 5))
    g = mx.nd.zeros((4, 5))
    exe = sym.bind(mx.cpu(), args={"data": x}, args_grad={"data": g}, aux_states={"sq_aux": mx.nd.zeros((4, 5))})
    exe.forward(is_train=True)
    exe.backward(mx.nd.ones((4, 5)))
    assert np.allclose(exe.outputs[0].asnumpy(), x.asnumpy() ** 2) and np.allclose(g.asnumpy(), 2 * x.asnumpy())
    assert (exe.aux_dict["sq_aux"].asnumpy() == 1).all()
    op = mx.operator.NumpyOp()
    s = op.get_symbol(mx.sym.var("X"), name="np_op")
    dx = mx.nd.zeros((10,))
    e = s.bind(mx.cpu(), args=[mx.nd.ones((10,)) * 10], args_grad={"X": dx})
    e.forward(is_train=True)
    e.backward(mx.nd.ones((10,)))
    assert (e.outputs[0].asnumpy() == 10).all() and (dx.asnumpy() == 1).all()


def test_linalg_axis_and_fp64_precision():
    rs = np.random.RandomState(4)
    a, b = rs.rand(3, 1, 2, 3), rs.rand(3, 1, 3, 2)
    r = np.matmul(a, b)
    a2, b2, r2 = np.swapaxes(a, 0, 2), np.swapaxes(b, 0, 2), np.swapaxes(r, 0, 2)
    got = mx.nd.linalg.gemm2(mx.nd.array(a2, dtype="float64"), mx.nd.array(b2, dtype="float64"), axis=0).asnumpy()
    assert np.allclose(got, r2, rtol=1e-12)
    m = rs.rand(4, 4)
    spd = m @ m.T + 4 * np.eye(4)
    L = np.linalg.cholesky(spd)
    assert np.allclose(mx.nd.linalg.potri(mx.nd.array(L, dtype="float64")).asnumpy(), np.linalg.inv(spd), rtol=1e-10)


def test_pooling_same_convention_depth_to_space_errors_tile_errors():
    x = mx.nd.array(np.random.RandomState(5).rand(1, 1, 9))
    out = mx.nd.Pooling(x, kernel=2, stride=2, pad=(0, 0, 0), pool_type="max", pooling_convention="same")
    assert out.shape == (1, 1, 5) and out.asnumpy()[0, 0, 4] == x.asnumpy()[0, 0, 8]
    with pytest.raises(MXNetError):
        mx.nd.depth_to_space(mx.nd.zeros((1, 3, 2, 2)), 2)
    with pytest.raises(MXNetError):
        mx.nd.space_to_depth(mx.nd.zeros((1, 4, 3, 2)), 2)
    with pytest.raises(MXNetError):
        mx.nd.tile(mx.nd.zeros((2, 2)), (1, 0, 3))


def test_partial_shape_inference_of_slices_and_reshape_backward():
    v = mx.sym.var("data", shape=(0, 20))
    assert mx.sym.slice(v, begin=(None, 3), end=(5, 10)).infer_shape_partial()[1] == [(0, 7)]
    assert mx.sym.slice_axis(mx.sym.var("d", shape=(10, 0)), axis=0, begin=0, end=5).infer_shape_partial()[1] == [(5, 0)]
    net = mx.sym.elemwise_add(mx.sym.var("x").reshape((0, -1)), mx.sym.ones(shape=(2, 75)))
    args, outs, _ = net.infer_shape(x=(2, 3, 0, 5))
    assert args == [(2, 3, 5, 5)] and outs == [(2, 75)]
    b = mx.sym.topk(mx.sym.var("a"), axis=None, k=3, ret_typ="value")
    assert b.eval(a=mx.nd.array(np.arange(24.0).reshape(2, 3, 4)))[0].asnumpy().tolist() == [23.0, 22.0, 21.0]


def test_samplers_unique_zipfian_gamma_zero_multinomial_bound():
    classes, tries = mx.nd._internal._sample_unique_zipfian(1000, shape=(2, 64))
    for i in range(2):
        row = classes[i].asnumpy()
        assert len(set(row.tolist())) == 64 and row.min() >= 0 and row.max() < 1000
        assert int(tries[i].asscalar()) >= 64
    g = mx.nd.random.gamma(mx.nd.array([0.0, 2.0]), mx.nd.array([1.0, 1.0]), shape=100).asnumpy()
    assert (g[0] == 0).all() and (g[1] > 0).all()
    with pytest.raises(MXNetError):
        mx.nd.random.multinomial(mx.nd.zeros(300), dtype="uint8")


def test_sparse_storage_rules_constructors_and_format_checks():
    sp = mx.nd.sparse
    dense = np.zeros((6, 3), "f")
    dense[[1, 4]] = [[1, -2, 3], [-4, 5, -6]]
    rsp = mx.nd.array(dense).tostype("row_sparse")
    assert mx.nd.relu(rsp).stype == "row_sparse" and np.allclose(mx.nd.relu(rsp).asnumpy(), np.maximum(dense, 0))
    assert mx.nd.elemwise_mul(mx.nd.ones((6, 3)), rsp).stype == "row_sparse"
    assert mx.nd.elemwise_add(mx.nd.ones((6, 3)), rsp).stype == "default"
    assert mx.nd.elemwise_add(rsp, rsp).stype == "row_sparse"
    assert sp.row_sparse_array((np.int32(4), np.int64(2))).shape == (4, 2)
    with pytest.raises(ValueError):
        sp.row_sparse_array((4, 2), shape=(5, 2))
    with pytest.raises(ValueError):
        sp.csr_matrix((mx.nd.zeros((0,)), mx.nd.zeros((0,)), mx.nd.zeros((3,))))      # shape cannot be inferred
    with pytest.raises(ValueError):
        sp.zeros("no_such_stype", (2, 2))
    bad = sp.csr_matrix(([7, 8, 9], [2, 1, 1], [0, 2, 2, 3]), shape=(3, 4))
    with pytest.raises(MXNetError):
        bad.check_format()
    with pytest.raises(MXNetError):
        sp.row_sparse_array(([[1, 2], [3, 4]], [1, 0]), shape=(3, 2)).check_format()
    import pickle
    back = pickle.loads(pickle.dumps(rsp))
    assert back.stype == "row_sparse" and np.allclose(back.asnumpy(), dense)
    out = mx.nd.zeros((6, 3)).tostype("row_sparse")
    mx.nd.array(dense).copyto(out)
    assert out.stype == "row_sparse" and np.allclose(out.asnumpy(), dense)
    d = sp.dot(mx.nd.array(dense.T), mx.nd.array(dense).tostype("csr"), forward_stype="csr")
    assert d.stype == "csr" and np.allclose(d.asnumpy(), dense.T @ dense)


def test_deferred_exception_is_delivered_once_and_reaches_executor_gradients():
    a = mx.nd.random.normal(0, -1, (2, 2))
    with pytest.raises(MXNetError):
        a.asnumpy()
    a.asnumpy()                                   # delivered once: later reads do not raise again
    x = mx.sym.var("x")
    out = mx.sym.make_loss(mx.sym.dot(x, mx.sym.random.normal(0, -1, (2, 2))))
    exe = out.bind(mx.cpu(), args={"x": mx.nd.ones((2, 2))}, args_grad={"x": mx.nd.zeros((2, 2))})
    exe.forward(is_train=True)
    exe.backward()
    with pytest.raises(MXNetError):
        exe.grad_arrays[0].asnumpy()


def test_maximum_minimum_gradient_goes_to_the_left_operand_at_ties():
    a, b = mx.nd.array([1.0, 2.0, 3.0]), mx.nd.array([1.0, 5.0, 0.0])
    a.attach_grad()
    b.attach_grad()
    with mx.autograd.record():
        y = mx.nd.maximum(a, b) + mx.nd.minimum(a, b)
    y.backward()
    assert y.asnumpy().tolist() == [2.0, 7.0, 3.0]
    assert a.grad.asnumpy().tolist() == [2.0, 1.0, 1.0] and b.grad.asnumpy().tolist() == [0.0, 1.0, 1.0]
