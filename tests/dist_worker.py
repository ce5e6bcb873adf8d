"""Worker used by tests/test_dist_kvstore.py (closed-form oracle of tests/nightly/dist_sync_kvstore.py:102-115)."""
import os
import sys

import faulthandler

import numpy as np

faulthandler.dump_traceback_later(90, exit=True)   # a hung collective must fail the test, not stall it

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mxb200 as mx  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "default"
# "gpu": worker-side arrays live on cuda:0 (dist_device_sync: device gradients staged through pinned host buffers)
ctx = mx.gpu(0) if mode == "gpu" else mx.cpu()
_nd = mx.nd


class _OnCtx(object):
    """mx.nd factory functions that allocate on ``ctx``."""

    def __getattr__(self, name):
        f = getattr(_nd, name)
        if name in ("ones", "zeros", "array"):
            return lambda *a, **k: f(*a, ctx=ctx, **k)
        return f


if mode == "gpu":
    mx.nd = _OnCtx()
kv = mx.kv.create("dist_device_sync" if mode == "gpu" else "dist_sync")
rank, n = kv.rank, kv.num_workers
shape = (3, 4)
rate = 2.0
kv.init("w", mx.nd.ones(shape))
kv.init(["big"], [mx.nd.zeros((1200, 1000))])          # > bigarray_bound: sharded over all servers
if mode == "compressed":
    kv.set_gradient_compression({"type": "2bit", "threshold": 0.5})
kv.set_optimizer(mx.optimizer.create("test", rescale_grad=rate))
out = mx.nd.zeros(shape)
if mode == "compressed":
    # every worker pushes 0.6 -> quantised to +0.5 with residual 0.1; second push 0.6 -> residual 0.7 -> +0.5 again
    for i in range(2):
        kv.push("w", mx.nd.ones(shape) * 0.6)
        kv.pull("w", out=out)
        assert np.allclose(out.asnumpy(), 1 + rate * 0.5 * n * (i + 1)), out.asnumpy()[0]
else:
    for i in range(3):
        kv.push("w", mx.nd.ones(shape) * (rank + 1))
        kv.pull("w", out=out)
        expect = (n + 1) * n * rate / 2 * (i + 1) + 1
        assert np.allclose(out.asnumpy(), expect), (out.asnumpy()[0], expect)
    big = mx.nd.zeros((1200, 1000))
    kv.push("big", mx.nd.ones((1200, 1000)) * (rank + 1))
    kv.pull("big", out=big)
    assert np.allclose(big.asnumpy(), (n + 1) * n * rate / 2), big.asnumpy()[0, :3]
    kv.init("aux", mx.nd.zeros((5,)), exclude_update=True)
    kv.push("aux", mx.nd.ones((5,)) * (rank + 1))
    a = mx.nd.zeros((5,))
    kv.pull("aux", out=a)
    assert np.allclose(a.asnumpy(), (n + 1) / 2.0), a.asnumpy()
    rs = mx.nd.sparse.zeros("row_sparse", shape)
    kv.row_sparse_pull("w", out=rs, row_ids=mx.nd.array([2, 0, 2]))
    assert rs.indices.asnumpy().tolist() == [0, 2]
    assert kv.get_num_dead_node() == 0
print("worker %d/%d OK" % (rank, n), flush=True)
