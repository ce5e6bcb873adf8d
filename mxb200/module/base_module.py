"""BaseModule: the high-level train / score / predict loop, including the elastic-training epoch loop
(reference: python/mxnet/module/base_module.py:412-623; elastic additions :503-623).

Provenance: adapted from Apache MXNet (incubating) 1.3.1 ``python/mxnet/module/base_module.py`` (Apache License 2.0).  The public API, argument
names and control flow follow the original so that user scripts keep working; the tensor maths runs on torch /
the mxb200 kernels.  This file is API-compatibility glue, not original B200 work ("Provenance").
"""
import logging
import os
import time
import warnings

import numpy as np

from .. import metric as _metric
from .. import ndarray as nd
from ..context import cpu
from ..model import BatchEndParam
from ..initializer import Uniform
from ..io import DataDesc, DataBatch, DataIter
from ..ndarray import NDArray
from ..base import getenv

__all__ = ["BaseModule"]


def _as_list(obj):
    return obj if isinstance(obj, list) else [obj]


def _check_input_names(symbol, names, typename, throw):
    args = symbol.list_arguments()
    for name in names:
        if name in args:
            continue
        candidates = [arg for arg in args if not arg.endswith("_weight") and not arg.endswith("_bias") and
                      not arg.endswith("_gamma") and not arg.endswith("_beta")]
        msg = "\033[91mYou created Module with Module(..., %s_names=%s) but input with name '%s' is not found in " \
              "symbol.list_arguments(). Did you mean one of:\n\t%s\033[0m" % (typename, str(names), name, "\n\t".join(candidates))
        if throw:
            raise ValueError(msg)
        warnings.warn(msg)


def _parse_data_desc(data_names, label_names, data_shapes, label_shapes):
    data_shapes = [x if isinstance(x, DataDesc) else DataDesc(*x) for x in data_shapes]
    _check_names_match(data_names, data_shapes, "data", True)
    if label_shapes is not None:
        label_shapes = [x if isinstance(x, DataDesc) else DataDesc(*x) for x in label_shapes]
        _check_names_match(label_names, label_shapes, "label", False)
    else:
        _check_names_match(label_names, [], "label", False)
    return data_shapes, label_shapes


def _check_names_match(data_names, data_shapes, name, throw):
    actual = [x[0] for x in data_shapes]
    if sorted(data_names) != sorted(actual):
        msg = "Data provided by %s_shapes don't match names specified by %s_names (%s vs. %s)" % (
            name, name, str(data_shapes), str(data_names))
        if throw:
            raise ValueError(msg)
        warnings.warn(msg)


class BaseModule(object):
    def __init__(self, logger=logging):
        self.logger = logger
        self.binded = False
        self.for_training = False
        self.inputs_need_grad = False
        self.params_initialized = False
        self.optimizer_initialized = False
        self._symbol = None
        self._total_exec_bytes = 0

    # ---------------------------------------------------------------------------------- high level
    def forward_backward(self, data_batch):
        self.forward(data_batch, is_train=True)
        self.backward()

    def score(self, eval_data, eval_metric, num_batch=None, batch_end_callback=None, score_end_callback=None, reset=True,
              epoch=0, sparse_row_id_fn=None):
        assert self.binded and self.params_initialized
        if reset:
            eval_data.reset()
        if not isinstance(eval_metric, _metric.EvalMetric):
            eval_metric = _metric.create(eval_metric)
        eval_metric.reset()
        actual_num_batch = 0
        for nbatch, eval_batch in enumerate(eval_data):
            if num_batch is not None and nbatch == num_batch:
                break
            self.prepare(eval_batch, sparse_row_id_fn=sparse_row_id_fn)
            self.forward(eval_batch, is_train=False)
            self.update_metric(eval_metric, eval_batch.label)
            if batch_end_callback is not None:
                batch_end_params = BatchEndParam(epoch=epoch, nbatch=nbatch, eval_metric=eval_metric, locals=locals())
                for callback in _as_list(batch_end_callback):
                    callback(batch_end_params)
            actual_num_batch += 1
        if score_end_callback:
            params = BatchEndParam(epoch=epoch, nbatch=actual_num_batch, eval_metric=eval_metric, locals=locals())
            for callback in _as_list(score_end_callback):
                callback(params)
        return eval_metric.get_name_value()

    def iter_predict(self, eval_data, num_batch=None, reset=True, sparse_row_id_fn=None):
        assert self.binded and self.params_initialized
        if reset:
            eval_data.reset()
        for nbatch, eval_batch in enumerate(eval_data):
            if num_batch is not None and nbatch == num_batch:
                break
            self.prepare(eval_batch, sparse_row_id_fn=sparse_row_id_fn)
            self.forward(eval_batch, is_train=False)
            pad = eval_batch.pad or 0
            outputs = [out[0:out.shape[0] - pad] for out in self.get_outputs()]
            yield (outputs, nbatch, eval_batch)

    def predict(self, eval_data, num_batch=None, merge_batches=True, reset=True, always_output_list=False,
                sparse_row_id_fn=None):
        assert self.binded and self.params_initialized
        if isinstance(eval_data, (NDArray, np.ndarray)):
            # a bare array is one data batch stream (base_module.py:330-340 of the reference)
            from ..io import NDArrayIter
            if isinstance(eval_data, np.ndarray):
                eval_data = nd.array(eval_data)
            self.forward(DataBatch([eval_data]), is_train=False)
            return self.get_outputs()[0]
        if not isinstance(eval_data, DataIter):
            raise ValueError("eval_data must be of type NDArray or DataIter")
        if reset:
            eval_data.reset()
        output_list = []
        for nbatch, eval_batch in enumerate(eval_data):
            if num_batch is not None and nbatch == num_batch:
                break
            self.prepare(eval_batch, sparse_row_id_fn=sparse_row_id_fn)
            self.forward(eval_batch, is_train=False)
            pad = eval_batch.pad or 0
            outputs = [out[0:out.shape[0] - pad].copy() for out in self.get_outputs()]
            output_list.append(outputs)
        if len(output_list) == 0:
            return output_list
        if merge_batches:
            num_outputs = len(output_list[0])
            for out in output_list:
                assert len(out) == num_outputs, "Cannot merge batches, as num of outputs is not the same in mini-batches. " \
                                                "Maybe bucketing is used?"
            output_list2 = [nd.concatenate([out[i] for out in output_list]) for i in range(num_outputs)]
            if num_outputs == 1 and not always_output_list:
                return output_list2[0]
            return output_list2
        return output_list

    def fit(self, train_data, eval_data=None, eval_metric="acc", epoch_end_callback=None, batch_end_callback=None,
            kvstore="local", optimizer="sgd", optimizer_params=(("learning_rate", 0.01),), eval_end_callback=None,
            epoch_begin_callback=None, eval_batch_end_callback=None, initializer=Uniform(0.01), arg_params=None, aux_params=None,
            allow_missing=False, force_rebind=False, force_init=False, begin_epoch=0, num_epoch=None,
            validation_metric=None, monitor=None, sparse_row_id_fn=None):
        """Train.  With ELASTIC_TRAINING_ENABLED=1 the loop runs the fork's protocol (base_module.py:503-623):

        * a worker started with NEW_WORKER=1 begins at EPOCH_BEGIN and pulls parameters + aux states from the
          store instead of initialising it (init_optimizer(initialize_from_kvstore=True));
        * at every epoch boundary (except a newcomer's first) all workers meet in
          ``kv._membership_change_barrier({'EPOCH_BEGIN': epoch})`` where the coordinator applies host-file
          changes (removals before additions);
        * if ``kv.num_workers`` changed, the data iterators are rebuilt through ``Module.get_iterator(kv)``
          (user ``BaseDataIterator``), so shards and per-worker batch follow the new membership;
        * at epoch end aux states (BN statistics) are averaged across devices and pushed under exclude-update
          keys so that newcomers can pull them.
        """
        assert num_epoch is not None, "please specify number of epochs"
        is_new_worker = os.environ.get("NEW_WORKER", "0") == "1"
        if is_new_worker and os.environ.get("EPOCH_BEGIN") is not None:
            begin_epoch = int(os.environ["EPOCH_BEGIN"])
        elastic = os.environ.get("ELASTIC_TRAINING_ENABLED", "0") in ("1", "true", "True")

        self.bind(data_shapes=train_data.provide_data, label_shapes=train_data.provide_label, for_training=True,
                  force_rebind=force_rebind)
        if monitor is not None:
            self.install_monitor(monitor)
        self.init_params(initializer=initializer, arg_params=arg_params, aux_params=aux_params,
                         allow_missing=allow_missing, force_init=force_init)
        self.init_optimizer(kvstore=kvstore, optimizer=optimizer, optimizer_params=optimizer_params,
                            initialize_from_kvstore=is_new_worker) if elastic or is_new_worker else \
            self.init_optimizer(kvstore=kvstore, optimizer=optimizer, optimizer_params=optimizer_params)

        if validation_metric is None:
            validation_metric = eval_metric
        if not isinstance(eval_metric, _metric.EvalMetric):
            eval_metric = _metric.create(eval_metric)

        kv = getattr(self, "_kvstore", None)
        for epoch in range(begin_epoch, num_epoch):
            if epoch_begin_callback is not None:           # fork addition (base_module.py:415,537-539): before the barrier
                for callback in _as_list(epoch_begin_callback):
                    callback(epoch, self.symbol, arg_params, aux_params)
            if elastic and kv is not None:
                prev_workers = kv.num_workers
                if not (is_new_worker and epoch == begin_epoch):
                    kv._membership_change_barrier({"EPOCH_BEGIN": str(epoch)})
                if kv.num_workers != prev_workers or (is_new_worker and epoch == begin_epoch):
                    self.logger.info("Epoch[%d] membership changed: %d -> %d workers (rank %d)", epoch, prev_workers,
                                     kv.num_workers, kv.rank)
                    it = self.get_iterator(kv)
                    if it is not None:
                        train_data, eval_data = it
                    self._on_membership_change(kv, prev_workers)
            tic = time.time()
            eval_metric.reset()
            nbatch = 0
            data_iter = iter(train_data)
            end_of_batch = False
            next_data_batch = next(data_iter)
            while not end_of_batch:
                data_batch = next_data_batch
                if monitor is not None:
                    monitor.tic()
                self.forward_backward(data_batch)
                self.update()
                if isinstance(data_batch, list):
                    self.update_metric(eval_metric, [db.label for db in data_batch], pre_sliced=True)
                else:
                    self.update_metric(eval_metric, data_batch.label)
                try:
                    next_data_batch = next(data_iter)
                    self.prepare(next_data_batch, sparse_row_id_fn=sparse_row_id_fn)
                except StopIteration:
                    end_of_batch = True
                if monitor is not None:
                    monitor.toc_print()
                if end_of_batch:
                    eval_name_vals = eval_metric.get_name_value()
                if batch_end_callback is not None:
                    batch_end_params = BatchEndParam(epoch=epoch, nbatch=nbatch, eval_metric=eval_metric, locals=locals())
                    for callback in _as_list(batch_end_callback):
                        callback(batch_end_params)
                nbatch += 1
            for name, val in eval_name_vals:
                self.logger.info("Epoch[%d] Train-%s=%f", epoch, name, val)
            toc = time.time()
            self.logger.info("Epoch[%d] Time cost=%.3f", epoch, (toc - tic))
            # sync aux params across devices, and (elastic) publish them on the store
            arg_params_, aux_params_ = self.get_params()
            self.set_params(arg_params_, aux_params_)
            if elastic and kv is not None:
                self.store_aux_params()
            if epoch_end_callback is not None:
                for callback in _as_list(epoch_end_callback):
                    callback(epoch, self.symbol, arg_params_, aux_params_)
            if eval_data:
                res = self.score(eval_data, validation_metric, score_end_callback=eval_end_callback,
                                 batch_end_callback=eval_batch_end_callback, epoch=epoch)
                for name, val in res:
                    self.logger.info("Epoch[%d] Validation-%s=%f", epoch, name, val)
            train_data.reset()
        # asynchronous epoch-end work pushed to the dependency engine (checkpoint writes) lands before fit() returns
        from .. import engine as _engine
        _engine.get().wait_all()

    def _on_membership_change(self, kv, prev_workers):
        """Re-apply the dynamic-mini-batch learning-rate rule (hook location in the reference:
        python/mxnet/module/base_module.py:540-549).  A no-op for the reference's constant-global-batch recipe; with a
        ``DynamicMiniBatchScheduler(per_worker_batch=...)`` the learning rate follows the new global batch and, for an abrupt
        change, the momentum buffers are rescaled (momentum correction)."""
        o = getattr(self, "_optimizer", None)
        sched = getattr(o, "lr_scheduler", None) if o is not None else None
        new_workers = kv.num_workers
        if sched is None or not hasattr(sched, "on_membership_change") or new_workers == prev_workers:
            return None
        per_worker = getattr(sched, "per_worker_batch", None)
        new_batch = per_worker * new_workers if per_worker else sched.base_batch
        sched.on_membership_change(new_batch, num_update=o.num_update)
        corr = sched.momentum_correction() if sched.smooth_steps == 0 else 1.0
        dt = getattr(self, "_fused", None)
        if corr != 1.0 and getattr(o, "momentum", 0.0):
            if dt is not None:
                dt.scale_state("mom", corr)
            else:
                upd = getattr(self, "_updater", None)
                for st in (upd.states.values() if upd is not None else []):
                    for x in (st if isinstance(st, (list, tuple)) else [st]):
                        if hasattr(x, "_t") and x._t.dtype.is_floating_point:
                            x._t.mul_(corr)
        self.logger.info("membership %d -> %d workers: global batch %s, lr factor %.4f, momentum correction %.4f",
                         prev_workers, new_workers, new_batch, sched.cur_factor if sched.smooth_steps == 0 else sched._to, corr)
        return {"global_batch": new_batch, "momentum_correction": corr}

    # ---------------------------------------------------------------------------------- symbol / params
    @property
    def data_names(self):
        raise NotImplementedError()

    @property
    def output_names(self):
        raise NotImplementedError()

    @property
    def data_shapes(self):
        raise NotImplementedError()

    @property
    def label_shapes(self):
        raise NotImplementedError()

    @property
    def output_shapes(self):
        raise NotImplementedError()

    @property
    def symbol(self):
        return self._symbol

    def get_params(self):
        raise NotImplementedError()

    def get_iterator(self, kv):
        return None

    def store_aux_params(self):
        pass

    def init_params(self, initializer=Uniform(0.01), arg_params=None, aux_params=None, allow_missing=False,
                    force_init=False, allow_extra=False):
        raise NotImplementedError()

    def set_params(self, arg_params, aux_params, allow_missing=False, force_init=True, allow_extra=False):
        self.init_params(initializer=None, arg_params=arg_params, aux_params=aux_params, allow_missing=allow_missing,
                         force_init=force_init, allow_extra=allow_extra)

    def save_params(self, fname):
        arg_params, aux_params = self.get_params()
        save_dict = {("arg:%s" % k): v.as_in_context(cpu()) for k, v in arg_params.items()}
        save_dict.update({("aux:%s" % k): v.as_in_context(cpu()) for k, v in aux_params.items()})
        nd.save(fname, save_dict)

    def load_params(self, fname):
        save_dict = nd.load(fname)
        arg_params, aux_params = {}, {}
        for k, value in save_dict.items():
            arg_type, name = k.split(":", 1)
            if arg_type == "arg":
                arg_params[name] = value
            elif arg_type == "aux":
                aux_params[name] = value
            else:
                raise ValueError("Invalid param file " + fname)
        self.set_params(arg_params, aux_params)

    def get_states(self, merge_multi_context=True):
        assert self.binded and self.params_initialized
        assert not merge_multi_context
        return []

    def set_states(self, states=None, value=None):
        assert self.binded and self.params_initialized
        assert not states and not value

    def install_monitor(self, mon):
        raise NotImplementedError()

    def prepare(self, data_batch, sparse_row_id_fn=None):
        if sparse_row_id_fn is not None:
            warnings.warn(UserWarning("sparse_row_id_fn is not invoked for BaseModule."))

    def forward(self, data_batch, is_train=None):
        raise NotImplementedError()

    def backward(self, out_grads=None):
        raise NotImplementedError()

    def get_outputs(self, merge_multi_context=True):
        raise NotImplementedError()

    def get_input_grads(self, merge_multi_context=True):
        raise NotImplementedError()

    def update(self):
        raise NotImplementedError()

    def update_metric(self, eval_metric, labels, pre_sliced=False):
        raise NotImplementedError()

    def bind(self, data_shapes, label_shapes=None, for_training=True, inputs_need_grad=False, force_rebind=False,
             shared_module=None, grad_req="write"):
        raise NotImplementedError()

    def init_optimizer(self, kvstore="local", optimizer="sgd", optimizer_params=(("learning_rate", 0.01),),
                       force_init=False):
        raise NotImplementedError()
