"""KVStore semantics (reference tests: tests/python/unittest/test_kvstore.py; multi-device is simulated with
mx.cpu(0..3) contexts exactly like the reference does)."""
import numpy as np
import pytest

import mxb200 as mx

SHAPE = (4, 4)
KEYS = [5, 7, 11]
STR_KEYS = ["b", "c", "d"]


def init_kv(kind="local", str_keys=False):
    kv = mx.kv.create(kind)
    kv.init(3 if not str_keys else "a", mx.nd.zeros(SHAPE))
    kv.init(STR_KEYS if str_keys else KEYS, [mx.nd.zeros(SHAPE)] * len(KEYS))
    return kv


def check(A, x):
    assert np.sum(np.abs((A - x).asnumpy())) == 0


@pytest.mark.parametrize("kind", ["local", "device"])
@pytest.mark.parametrize("str_keys", [False, True])
def test_single_kv_pair_and_list(kind, str_keys):
    kv = init_kv(kind, str_keys)
    k = "a" if str_keys else 3
    kv.push(k, mx.nd.ones(SHAPE))
    val = mx.nd.empty(SHAPE)
    kv.pull(k, out=val)
    check(val, 1)
    keys = STR_KEYS if str_keys else KEYS
    kv.push(keys, [mx.nd.ones(SHAPE) * 4] * len(keys))
    vals = [mx.nd.empty(SHAPE)] * len(keys)
    kv.pull(keys, out=vals)
    for v in vals:
        check(v, 4)


@pytest.mark.parametrize("kind", ["local", "device"])
def test_aggregator_over_four_devices(kind):
    kv = init_kv(kind)
    num_devs = 4
    devs = [mx.cpu(i) for i in range(num_devs)]
    vals = [mx.nd.ones(SHAPE, d) for d in devs]
    kv.push(3, vals)
    kv.pull(3, out=vals)
    for v in vals:
        check(v, num_devs)
    vals = [[mx.nd.ones(SHAPE, d) * 2.0 for d in devs]] * len(KEYS)
    kv.push(KEYS, vals)
    kv.pull(KEYS, out=vals)
    for vv in vals:
        for v in vv:
            check(v, num_devs * 2.0)


def test_duplicate_keys_in_one_push_are_summed():
    kv = init_kv()
    kv.push([3, 3], [mx.nd.ones(SHAPE), mx.nd.ones(SHAPE) * 2])
    out = mx.nd.empty(SHAPE)
    kv.pull(3, out=out)
    check(out, 3)


def test_updater_and_optimizer():
    kv = init_kv()

    def updater(key, recv, local):
        local += recv * 2
    kv._set_updater(updater)
    devs = [mx.cpu(i) for i in range(4)]
    for _ in range(3):
        kv.push(3, [mx.nd.ones(SHAPE, d) for d in devs])
    out = mx.nd.empty(SHAPE)
    kv.pull(3, out=out)
    check(out, 3 * 4 * 2)
    kv2 = init_kv()
    kv2.set_optimizer(mx.optimizer.create("test", rescale_grad=0.5))
    kv2.push(3, mx.nd.ones(SHAPE))
    kv2.pull(3, out=out)
    check(out, 0.5)


def test_closed_form_sync_rounds():
    """dist_sync oracle of the reference (tests/nightly/dist_sync_kvstore.py:102-115) on a local store:
    n workers push (r+1)*ones each round with the Test optimizer (rate)."""
    n, rate = 4, 2.0
    kv = mx.kv.create("local")
    kv.init(9, mx.nd.ones(SHAPE))
    kv.set_optimizer(mx.optimizer.create("test", rescale_grad=rate))
    out = mx.nd.empty(SHAPE)
    for i in range(3):
        kv.push(9, [mx.nd.ones(SHAPE, mx.cpu(r)) * (r + 1) for r in range(n)])
        kv.pull(9, out=out)
        check(out, (n + 1) * n * rate / 2 * (i + 1) + 1)


def test_errors():
    kv = init_kv()
    with pytest.raises(mx.MXNetError):
        kv.init(3, mx.nd.ones(SHAPE))                       # double init
    with pytest.raises(mx.MXNetError):
        kv.push(99, mx.nd.ones(SHAPE))                      # unknown key
    with pytest.raises(mx.MXNetError):
        kv.pull("strkey", out=mx.nd.ones(SHAPE))           # mixing key types
    with pytest.raises(mx.MXNetError):
        kv._membership_change_barrier({"EPOCH_BEGIN": "1"})  # only for dist stores
    with pytest.raises(mx.MXNetError):
        mx.kv.create("local").set_gradient_compression({"type": "2bit"})
    with pytest.raises(mx.MXNetError):
        mx.kv.create("bogus")


def test_exclude_update_keys_are_averaged_not_optimised():
    kv = mx.kv.create("local")
    kv.set_optimizer(mx.optimizer.create("test", rescale_grad=10.0))
    kv.init("bn_moving_mean", mx.nd.zeros((3,)), exclude_update=True)
    kv.init("fc_weight", mx.nd.zeros((3,)))
    assert kv._str2int["bn_moving_mean"] == mx.kv.MAX_ALLOWED_KEY_FOR_UPDATE + 1
    kv.push("bn_moving_mean", mx.nd.ones((3,)) * 4)
    kv.push("fc_weight", mx.nd.ones((3,)))
    a, b = mx.nd.zeros((3,)), mx.nd.zeros((3,))
    kv.pull("bn_moving_mean", out=a)
    kv.pull("fc_weight", out=b)
    check(a, 4)         # stored value := merged / num_workers (1 here), optimizer NOT applied
    check(b, 10)


def test_row_sparse_pull():
    kv = mx.kv.create("local")
    w = np.arange(40, dtype=np.float32).reshape(10, 4)
    kv.init("e", mx.nd.array(w).tostype("row_sparse"))
    out = mx.nd.sparse.zeros("row_sparse", (10, 4))
    kv.row_sparse_pull("e", out=out, row_ids=mx.nd.array([7, 2, 2, 5]))
    assert out.indices.asnumpy().tolist() == [2, 5, 7]
    np.testing.assert_array_equal(out.data.asnumpy(), w[[2, 5, 7]])
    dense = mx.nd.zeros((10, 4))
    kv.pull("e", out=dense)     # ignored for sparse stored value
    assert dense.asnumpy().sum() == 0


def _np_2bit(grad, residual, threshold):
    """NumPy oracle (the reference's compute_expected_2bit_quantization, tests/nightly/test_kvstore.py:33-77)."""
    residual = residual + grad
    out = np.zeros_like(grad)
    out[residual >= threshold] = threshold
    out[residual <= -threshold] = -threshold
    residual = residual - out
    return out, residual


def test_two_bit_compression_matches_numpy_oracle():
    kv = mx.kv.create("device")
    kv.set_gradient_compression({"type": "2bit", "threshold": 0.5})
    shape = (3, 11)
    kv.init(1, mx.nd.zeros(shape))
    rng = np.random.RandomState(0)
    res = [np.zeros(shape, np.float32) for _ in range(2)]
    for _ in range(4):
        grads = [rng.uniform(-1, 1, shape).astype(np.float32) for _ in range(2)]
        expect = np.zeros(shape, np.float32)
        for i in range(2):
            q, res[i] = _np_2bit(grads[i], res[i], 0.5)
            expect += q
        kv.push(1, [mx.nd.array(g, ctx=mx.cpu(i)) for i, g in enumerate(grads)])
        out = mx.nd.zeros(shape)
        kv.pull(1, out=out)
        np.testing.assert_allclose(out.asnumpy(), expect)
    gc = kv._compression
    import torch
    words = gc.quantize(torch.tensor([0.6, -0.7, 0.1, 0.0]), torch.zeros(4))
    assert gc.compressed_size(4) == 1 and (int(words[0]) & 0xFF) == 0xE0   # 11 10 00 00


def test_gpu_topology_trees_and_tree_comm():
    """Reference: tests/cpp/kvstore/gpu_topology_test.cc checks that every generated tree is a valid spanning tree;
    plus a functional reduce/broadcast along the trees (on CPU tensors)."""
    import numpy as np
    import torch
    from mxb200.kvstore import gpu_topology as T
    # hybrid cube-mesh like DGX-1V (entries = number of NVLinks), a PCIe-only box and an asymmetric 6-GPU case
    dgx1 = np.array([[0, 1, 1, 2, 2, 0, 0, 0], [1, 0, 2, 1, 0, 2, 0, 0], [1, 2, 0, 2, 0, 0, 1, 0], [2, 1, 2, 0, 0, 0, 0, 1],
                     [2, 0, 0, 0, 0, 1, 1, 2], [0, 2, 0, 0, 1, 0, 2, 1], [0, 0, 1, 0, 1, 2, 0, 2], [0, 0, 0, 1, 2, 1, 2, 0]],
                    dtype=np.float32)
    flat = np.ones((4, 4), dtype=np.float32) - np.eye(4, dtype=np.float32)
    rng = np.random.RandomState(0)
    odd = rng.randint(0, 3, size=(6, 6)).astype(np.float32)
    for W in (dgx1, flat, odd):
        n = W.shape[0]
        trees = T.compute_trees(W)
        assert len(trees) == n
        for root, levels in enumerate(trees):
            edges = [e for lvl in levels for e in lvl]
            assert len(edges) == n - 1
            reached = {root}
            for lvl in levels:                      # a parent must already be in the tree when its level runs
                for p, c in lvl:
                    assert p in reached and c not in reached
                for p, c in lvl:
                    reached.add(c)
            assert reached == set(range(n))
            assert len(levels) <= int(np.ceil(np.log2(n))) + 1
        if W is dgx1:   # every hop of every tree rides an NVLink
            Ws = np.maximum(W, W.T)
            assert all(Ws[p, c] > 0 for levels in trees for lvl in levels for p, c in lvl)
    tc = T.TreeComm(list(range(8)), W=dgx1)
    xs = [torch.full((5,), float(i + 1)) for i in range(8)]
    for root in (0, 3, 7):
        assert torch.equal(tc.reduce(xs, root), torch.full((5,), 36.0))
        outs = [torch.zeros(5) for _ in range(8)]
        tc.broadcast(torch.full((5,), 2.5), outs, root)
        assert all(torch.equal(o, torch.full((5,), 2.5)) for o in outs)
    assert all(float(x[0]) == i + 1 for i, x in enumerate(xs))   # inputs untouched


def test_create_kvstore_joins_torchrun_job_before_deciding(tmp_path):
    """under torchrun with one device per process, ``_create_kvstore('device', 1, ...)`` must join the
    process group first -- otherwise every rank trains an unsynchronised replica with kv=None."""
    import os
    import subprocess
    import sys
    import textwrap
    code = textwrap.dedent('''
        import os, sys
        sys.path.insert(0, %r)
        os.environ.update(WORLD_SIZE="2", RANK=sys.argv[1], LOCAL_RANK=sys.argv[1], MASTER_ADDR="127.0.0.1", MASTER_PORT="29671")
        from mxb200.model import _create_kvstore
        kv, upd = _create_kvstore("device", 1, {})
        print("KV", type(kv).__name__, kv.num_workers if kv is not None else None)
    ''' % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    ps = [subprocess.Popen([sys.executable, "-c", code, str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
          for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in ps]
    for o in outs:
        assert "KV KVStoreDevice 2" in o, o
