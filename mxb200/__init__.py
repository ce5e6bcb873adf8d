"""mxb200 -- a Blackwell-native elastic data-parallel training framework with the capabilities and user
API of awslabs/dynamic-training-with-apache-mxnet-on-aws (an MXNet 1.3.1 fork).

    import mxb200 as mx
    kv  = mx.kv.create('device')
    net = mx.gluon.model_zoo.vision.resnet50_v1b()
    trainer = mx.gluon.Trainer(net.collect_params(), 'nag', {...}, kvstore=kv)

PyTorch supplies tensors/autograd; the hot ops, the fused allreduce+optimizer and the runtime pieces are
hand-written CUDA (sm_100a) and C++ under ``mxb200/csrc``.
"""
import importlib as _importlib

__version__ = "0.1.0"

from .base import MXNetError, bfloat16  # noqa: F401
from .context import Context, cpu, gpu, cpu_pinned, current_context, num_gpus  # noqa: F401
from . import base, context, autograd, engine  # noqa: F401
from . import ndarray  # noqa: F401
from . import ndarray as nd  # noqa: F401

_LAZY = {
    "optimizer": ".optimizer", "lr_scheduler": ".lr_scheduler", "initializer": ".initializer", "init": ".initializer",
    "kvstore": ".kvstore", "kv": ".kvstore", "metric": ".metric", "callback": ".callback", "io": ".io",
    "model": ".model", "module": ".module", "mod": ".module", "gluon": ".gluon", "symbol": ".symbol", "sym": ".symbol",
    "profiler": ".profiler", "monitor": ".monitor", "mon": ".monitor", "random": ".random", "rnd": ".random",
    "recordio": ".recordio", "image": ".image", "img": ".image", "test_utils": ".test_utils", "storage": ".storage",
    "parallel": ".parallel", "models": ".models", "ops": ".ops", "executor": ".executor", "rtc": ".rtc",
    "kvstore_server": ".kvstore_server", "contrib": ".contrib", "visualization": ".visualization", "viz": ".visualization",
    "name": ".name", "attribute": ".attribute", "operator": ".operator", "rnn": ".rnn", "log": ".log",
    "libinfo": ".libinfo", "util": ".util", "registry": ".registry", "misc": ".misc", "notebook": ".notebook",
    "executor_manager": ".executor_manager", "torch": ".torch",
}


def __getattr__(name):
    if name == "AttrScope":
        from .attribute import AttrScope
        return AttrScope
    target = _LAZY.get(name)
    if target is None:
        raise AttributeError("module 'mxb200' has no attribute %r" % name)
    mod = _importlib.import_module(target, __name__)
    globals()[name] = mod
    return mod


def __dir__():
    return sorted(list(globals().keys()) + list(_LAZY.keys()))
