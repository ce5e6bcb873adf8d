"""Step-by-step comparison of Module.update(): fused kernel vs per-key updater (debug aid)."""
import os
import sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mxb200 as mx  # noqa: E402

rs = np.random.RandomState(0)
x = rs.randn(64, 20).astype("f")
y = (x[:, 0] > 0).astype("f")
net = mx.sym.FullyConnected(mx.sym.var("data"), num_hidden=32, name="fc1")
net = mx.sym.Activation(net, act_type="relu")
net = mx.sym.SoftmaxOutput(mx.sym.FullyConnected(net, num_hidden=2, name="fc2"), name="softmax")
init = {"fc1_weight": rs.randn(32, 20).astype("f") * 0.1, "fc1_bias": rs.randn(32).astype("f") * 0.1,
        "fc2_weight": rs.randn(2, 32).astype("f") * 0.1, "fc2_bias": rs.randn(2).astype("f") * 0.1}
for fused in ("1", "0"):
    os.environ["MXB200_FUSED_MODULE"] = fused
    it = mx.io.NDArrayIter(x, y, batch_size=16)
    mod = mx.mod.Module(net, context=[mx.gpu(0)])
    mod.bind(it.provide_data, it.provide_label)
    mod.init_params(arg_params={k: mx.nd.array(v) for k, v in init.items()}, aux_params={})
    mod.init_optimizer(kvstore="device", optimizer="sgd", optimizer_params={"learning_rate": 0.05, "momentum": 0.9, "wd": 1e-3})
    ex = mod._exec_group.execs[0]
    print("fused", fused, getattr(mod, "_fused", None) is not None, "rescale", mod._optimizer.rescale_grad)
    step = 0
    for _ in range(3):
        it.reset()
        for batch in it:
            mod.forward_backward(batch)
            g = ex.grad_dict["fc1_weight"]._t
            w = ex.arg_dict["fc1_weight"]._t
            gs = {k: float(ex.grad_dict[k]._t.abs().sum()) for k in init}
            mod.update()
            print(" step %2d  w00 %.8f  b0 %.8f  w2 %.8f  gsum %s  lr %s nu %s" % (
                step, float(ex.arg_dict["fc1_weight"]._t[0, 0]), float(ex.arg_dict["fc1_bias"]._t[0]),
                float(ex.arg_dict["fc2_weight"]._t[0, 0]), {k: round(v, 4) for k, v in gs.items()},
                mod._optimizer.lr, mod._optimizer.num_update))
            step += 1
    print(" final get_params w00 %.8f" % float(mod.get_params()[0]["fc1_weight"].asnumpy()[0, 0]))
