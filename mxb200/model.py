"""Model helpers: kvstore creation/initialisation, parameter update helpers, checkpoints
(reference: python/mxnet/model.py:77-227,434-492)."""
import logging
import os
from collections import namedtuple

import numpy as np

from . import kvstore as kvs
from . import ndarray as nd
from .ndarray import NDArray
from .base import getenv

BatchEndParam = namedtuple("BatchEndParams", ["epoch", "nbatch", "eval_metric", "locals"])


def _create_kvstore(kvstore, num_device, arg_params):
    """Returns (kv, update_on_kvstore) (model.py:77-114): no store for a single device unless dist;
    'local' with a parameter > 16M elements updates locally."""
    update_on_kvstore = True
    if kvstore is None:
        kv = None
    elif isinstance(kvstore, kvs.KVStore):
        kv = kvstore
    elif isinstance(kvstore, str):
        import torch.distributed as dist
        if ("device" in kvstore or "nccl" in kvstore) and "dist" not in kvstore:
            # a script started by torchrun joins the process group lazily: do it BEFORE deciding whether this is a
            # multi-process job, otherwise every rank would train an unsynchronised replica with no store at all
            kvs.kvstore._maybe_join_torchrun_job(dist)
        multi_proc = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        if num_device == 1 and "dist" not in kvstore and not multi_proc:
            kv = None
        else:
            kv = kvs.create(kvstore)
            if kvstore == "local":
                max_size = max([np.prod(param.shape) for param in arg_params.values()] or [0])
                if max_size > 1024 * 1024 * 16:
                    update_on_kvstore = False
    else:
        raise TypeError("kvstore must be KVStore, str or None")
    if kv is None:
        update_on_kvstore = False
    return (kv, update_on_kvstore)


def _initialize_kvstore(kvstore, param_arrays, arg_params, param_names, update_on_kvstore, aux_arrays=None,
                        aux_params=None, aux_names=None, exclude_update_aux=True):
    """kv.init for every parameter (and, for elastic training, every aux state under exclude-update keys) then
    pull into the device copies (model.py:116-134)."""
    for idx, param_on_devs in enumerate(param_arrays):
        name = param_names[idx]
        kvstore.init(name, arg_params[name])
        if update_on_kvstore:
            kvstore.pull(name, param_on_devs, priority=-idx)
    if aux_arrays is not None:
        for idx, aux_on_devs in enumerate(aux_arrays):
            name = aux_names[idx]
            kvstore.init(name, aux_params[name], exclude_update_aux)
            kvstore.pull(name, aux_on_devs, priority=-idx)


def _store_aux_params_on_kvstore(aux_arrays, kvstore, aux_names):
    """Push aux states (BN running stats): exclude-update keys are averaged over workers (model.py:175-181)."""
    for index, aux_list in enumerate(aux_arrays):
        name = aux_names[index]
        kvstore.push(name, aux_list[0] if len(aux_list) == 1 else aux_list, priority=-index)
        kvstore.pull(name, aux_list, priority=-index)


def _update_params_on_kvstore(param_arrays, grad_arrays, kvstore, param_names, aux_arrays=None, aux_names=None):
    """push every gradient, then pull every weight (model.py:136-152).  The reference interleaves push / pull per key and
    relies on its engine to run them asynchronously; issuing all pushes first is the same dataflow and lets the device
    store reduce ALL keys in one bucketed launch of the fused kernel (kvstore/fused_local.py)."""
    live = [(index, pair) for index, pair in enumerate(zip(param_arrays, grad_arrays)) if pair[1][0] is not None]
    for index, (arg_list, grad_list) in live:
        kvstore.push(param_names[index], grad_list, priority=-index)
    for index, (arg_list, grad_list) in live:
        kvstore.pull(param_names[index], arg_list, priority=-index)


def _update_params_on_kvstore_nccl(param_arrays, grad_arrays, kvstore, param_names):
    """Aggregated push/pull (MXNET_UPDATE_AGGREGATION_SIZE keys per call, model.py:155-173)."""
    from .base import getenv
    valid = [i for i, g in enumerate(grad_arrays) if g[0] is not None]
    size = getenv("MXNET_UPDATE_AGGREGATION_SIZE", 16)
    start = 0
    while start < len(valid):
        idx = valid[start:start + size]
        names = [param_names[i] for i in idx]
        kvstore.push(names, [grad_arrays[i] for i in idx], priority=-start)
        kvstore.pull(names, [param_arrays[i] for i in idx], priority=-start)
        start += size


def _update_params(param_arrays, grad_arrays, updater, num_device, kvstore=None, param_names=None):
    live = [(i, pair) for i, pair in enumerate(zip(param_arrays, grad_arrays)) if pair[1][0] is not None]
    if kvstore:
        # all pushes, then all pulls: one bucketed reduction for the whole parameter set (see _update_params_on_kvstore)
        for index, (arg_list, grad_list) in live:
            kvstore.push(param_names[index], grad_list, priority=-index)
        for index, (arg_list, grad_list) in live:
            kvstore.pull(param_names[index], grad_list, priority=-index)
    for index, (arg_list, grad_list) in live:
        for k, p in enumerate(zip(arg_list, grad_list)):
            w, g = p
            updater(index * num_device + k, g, w)


def save_checkpoint(prefix, epoch, symbol, arg_params, aux_params, async_write=False):
    """prefix-symbol.json + prefix-%04d.params with arg:/aux: keys (model.py:434-460)."""
    if symbol is not None:
        symbol.save("%s-symbol.json" % prefix)
    save_dict = {("arg:%s" % k): v for k, v in arg_params.items()}
    save_dict.update({("aux:%s" % k): v for k, v in aux_params.items()})
    param_name = "%s-%04d.params" % (prefix, epoch)
    # async_write (used by the epoch-end checkpoint callbacks): the arrays are snapshotted now, the file is written by the
    # dependency engine while the next epoch trains; readers of the file (nd.load / load_checkpoint), mx.nd.waitall(), the
    # end of fit() and the elastic re-init drain wait for it
    nd.save(param_name, save_dict, async_write=bool(async_write))
    logging.info('Saved checkpoint to "%s"', param_name)


def load_checkpoint(prefix, epoch):
    """Returns (symbol, arg_params, aux_params) (model.py:462-492)."""
    from . import symbol as sym
    symbol = sym.load("%s-symbol.json" % prefix) if os.path.exists("%s-symbol.json" % prefix) else None
    save_dict = nd.load("%s-%04d.params" % (prefix, epoch))
    arg_params, aux_params = {}, {}
    for k, v in save_dict.items():
        tp, name = k.split(":", 1)
        if tp == "arg":
            arg_params[name] = v
        if tp == "aux":
            aux_params[name] = v
    return (symbol, arg_params, aux_params)


class FeedForward(object):
    """Legacy model API (model.py:494-1063) kept as a thin adapter over Module."""

    def __init__(self, symbol, ctx=None, num_epoch=None, epoch_size=None, optimizer="sgd", initializer=None,
                 numpy_batch_size=128, arg_params=None, aux_params=None, allow_extra_params=False, begin_epoch=0,
                 **kwargs):
        from .initializer import Uniform
        self.symbol, self.ctx = symbol, ctx
        self.num_epoch, self.epoch_size, self.optimizer = num_epoch, epoch_size, optimizer
        self.initializer = initializer or Uniform(0.01)
        self.numpy_batch_size = numpy_batch_size
        self.arg_params, self.aux_params = arg_params, aux_params
        self.begin_epoch = begin_epoch
        self.kwargs = kwargs
        self._mod = None

    def _module(self, data):
        from .module import Module
        if self._mod is None:
            label_names = [d[0] if isinstance(d, tuple) else d.name for d in (data.provide_label or [])]
            self._mod = Module(self.symbol, data_names=[d[0] if isinstance(d, tuple) else d.name for d in data.provide_data],
                               label_names=label_names, context=self.ctx or [])
        return self._mod

    def fit(self, X, y=None, eval_data=None, eval_metric="acc", epoch_end_callback=None, batch_end_callback=None,
            kvstore="local", logger=None, **kw):
        data = self._as_iter(X, y)
        mod = self._module(data)
        mod.fit(data, eval_data=eval_data, eval_metric=eval_metric, epoch_end_callback=epoch_end_callback,
                batch_end_callback=batch_end_callback, kvstore=kvstore, optimizer=self.optimizer,
                optimizer_params=self.kwargs or (("learning_rate", 0.01),), initializer=self.initializer,
                arg_params=self.arg_params, aux_params=self.aux_params, begin_epoch=self.begin_epoch,
                num_epoch=self.num_epoch)
        self.arg_params, self.aux_params = mod.get_params()

    def _as_iter(self, X, y):
        from .io import DataIter, NDArrayIter
        if isinstance(X, DataIter):
            return X
        return NDArrayIter(X, y, batch_size=min(self.numpy_batch_size, len(X)), shuffle=False)

    def predict(self, X, num_batch=None):
        data = self._as_iter(X, None)
        mod = self._module(data)
        if not mod.binded:
            mod.bind(data.provide_data, None, for_training=False)
            mod.set_params(self.arg_params, self.aux_params)
        return mod.predict(data, num_batch=num_batch).asnumpy()

    def save(self, prefix, epoch=None):
        save_checkpoint(prefix, epoch if epoch is not None else self.num_epoch, self.symbol, self.arg_params, self.aux_params)

    @staticmethod
    def load(prefix, epoch, ctx=None, **kwargs):
        symbol, arg_params, aux_params = load_checkpoint(prefix, epoch)
        return FeedForward(symbol, ctx=ctx, arg_params=arg_params, aux_params=aux_params, begin_epoch=epoch, **kwargs)
