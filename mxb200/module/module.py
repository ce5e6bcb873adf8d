"""Module: symbolic training module over one or more contexts (reference: python/mxnet/module/module.py).

Fork-specific surface kept: ``Module(data_iterator=BaseDataIterator)`` + ``get_iterator(kv)``
(module.py:72-75,198-202), ``init_optimizer(initialize_from_kvstore=...)`` for workers joining a running job
(:485-571), ``store_aux_params`` (:668-683) and aux arrays registered on the kvstore under exclude-update keys
(model.py:116-134).

Provenance: adapted from Apache MXNet (incubating) 1.3.1 ``python/mxnet/module/module.py`` (Apache License 2.0).  The public API, argument
names and control flow follow the original so that user scripts keep working; the tensor maths runs on torch /
the mxb200 kernels.  This file is API-compatibility glue, not original B200 work ("Provenance").
"""
import logging
import warnings

import numpy as np
import torch

from .. import context as ctx_mod
from .. import optimizer as opt
from .. import ndarray as nd
from ..ndarray import NDArray
from ..initializer import Uniform, InitDesc
from ..io import DataDesc
from ..model import (_create_kvstore, _initialize_kvstore, _update_params, _update_params_on_kvstore,
                     _store_aux_params_on_kvstore, load_checkpoint, save_checkpoint)
from .base_module import BaseModule, _check_input_names, _parse_data_desc
from .executor_group import DataParallelExecutorGroup

__all__ = ["Module"]


class Module(BaseModule):
    def __init__(self, symbol, data_names=("data",), label_names=("softmax_label",), logger=logging,
                 context=None, work_load_list=None, fixed_param_names=None, state_names=None, group2ctxs=None,
                 compression_params=None, data_iterator=None):
        super(Module, self).__init__(logger=logger)
        if context is None or context == []:
            context = ctx_mod.cpu()
        if isinstance(context, ctx_mod.Context):
            context = [context]
        self._context = context
        if work_load_list is None:
            work_load_list = [1] * len(self._context)
        assert len(work_load_list) == len(self._context)
        self._work_load_list = work_load_list
        self._group2ctxs = group2ctxs
        self._symbol = symbol
        self._data_iterator = data_iterator
        data_names = list(data_names) if data_names is not None else []
        label_names = list(label_names) if label_names is not None else []
        state_names = list(state_names) if state_names is not None else []
        fixed_param_names = list(fixed_param_names) if fixed_param_names is not None else []
        _check_input_names(symbol, data_names, "data", True)
        _check_input_names(symbol, label_names, "label", False)
        _check_input_names(symbol, state_names, "state", True)
        _check_input_names(symbol, fixed_param_names, "fixed_param", True)
        arg_names = symbol.list_arguments()
        input_names = data_names + label_names + state_names
        self._param_names = [x for x in arg_names if x not in input_names]
        self._fixed_param_names = fixed_param_names
        self._aux_names = symbol.list_auxiliary_states()
        self._data_names = data_names
        self._label_names = label_names
        self._state_names = state_names
        self._output_names = symbol.list_outputs()
        self._arg_params = None
        self._aux_params = None
        self._params_dirty = False
        self._compression_params = compression_params
        self._optimizer = None
        self._kvstore = None
        self._update_on_kvstore = None
        self._updater = None
        self._preload_opt_states = None
        self._grad_req = None
        self._exec_group = None
        self._data_shapes = None
        self._label_shapes = None

    def get_iterator(self, kv):
        """(train_iter, val_iter) for the current membership, from the user's BaseDataIterator."""
        if self._data_iterator is None:
            return None
        return self._data_iterator.get_data_iterator(kv)

    @staticmethod
    def load(prefix, epoch, load_optimizer_states=False, **kwargs):
        sym, args, auxs = load_checkpoint(prefix, epoch)
        mod = Module(symbol=sym, **kwargs)
        mod._arg_params = args
        mod._aux_params = auxs
        mod.params_initialized = True
        if load_optimizer_states:
            mod._preload_opt_states = "%s-%04d.states" % (prefix, epoch)
        return mod

    def save_checkpoint(self, prefix, epoch, save_optimizer_states=False):
        self._symbol.save("%s-symbol.json" % prefix)
        param_name = "%s-%04d.params" % (prefix, epoch)
        self.save_params(param_name)
        logging.info('Saved checkpoint to "%s"', param_name)
        if save_optimizer_states:
            state_name = "%s-%04d.states" % (prefix, epoch)
            self.save_optimizer_states(state_name)
            logging.info('Saved optimizer state to "%s"', state_name)

    def _reset_bind(self):
        self._drop_fused()
        self.binded = False
        self._exec_group = None
        self._data_shapes = None
        self._label_shapes = None

    @property
    def data_names(self):
        return self._data_names

    @property
    def label_names(self):
        return self._label_names

    @property
    def output_names(self):
        return self._output_names

    @property
    def data_shapes(self):
        assert self.binded
        return self._data_shapes

    @property
    def label_shapes(self):
        assert self.binded
        return self._label_shapes

    @property
    def output_shapes(self):
        assert self.binded
        return self._exec_group.get_output_shapes()

    def get_params(self):
        assert self.binded and self.params_initialized
        if self._params_dirty:
            self._sync_params_from_devices()
        return (self._arg_params, self._aux_params)

    def init_params(self, initializer=Uniform(0.01), arg_params=None, aux_params=None, allow_missing=False,
                    force_init=False, allow_extra=False):
        if self.params_initialized and not force_init:
            warnings.warn("Parameters already initialized and force_init=False. init_params call ignored.", stacklevel=2)
            return
        assert self.binded, "call bind before initializing the parameters"

        def _impl(name, arr, cache):
            if cache is not None:
                if name in cache:
                    cache_arr = cache[name]
                    if cache_arr is not arr:
                        if getattr(arr, "stype", "default") != "default":
                            arr[:] = cache_arr                      # sparse host copy: storage-aware assignment
                        else:
                            with torch.no_grad():
                                arr._t.copy_(cache_arr._t.to(arr._t.dtype))
                else:
                    if not allow_missing:
                        raise RuntimeError("%s is not presented" % name)
                    if initializer is not None:
                        initializer(InitDesc(name, attrs.get(name, None)), arr)
            else:
                initializer(InitDesc(name, attrs.get(name, None)), arr)

        attrs = self._symbol.attr_dict()
        if self._arg_params is None:
            self._arg_params = {name: nd.zeros(arrs[0].shape, dtype=arrs[0].dtype)
                                for name, arrs in zip(self._param_names, self._exec_group.param_arrays)}
        if self._aux_params is None:
            self._aux_params = {name: nd.zeros(arrs[0].shape, dtype=arrs[0].dtype)
                                for name, arrs in zip(self._aux_names, self._exec_group.aux_arrays)}
        for name, arr in sorted(self._arg_params.items()):
            _impl(name, arr, arg_params)
        for name, arr in sorted(self._aux_params.items()):
            _impl(name, arr, aux_params)
        # variables declared with a sparse storage type keep it in the host-side copy (module.py:init_params allocates
        # ``_arg_params`` with the executor arrays' stype); the executor's working arrays stay dense
        for name in list(self._arg_params):
            st = (attrs.get(name) or {}).get("__storage_type__")
            st = {"1": "row_sparse", "2": "csr"}.get(str(st), st)
            if st in ("row_sparse", "csr") and getattr(self._arg_params[name], "stype", "default") == "default":
                self._arg_params[name] = self._arg_params[name].tostype(st)
        self.params_initialized = True
        self._params_dirty = False
        self._exec_group.set_params(self._arg_params, self._aux_params, allow_extra=allow_extra)
        if getattr(self, "_fused", None) is not None:
            self._fused.refresh_master()

    def set_params(self, arg_params, aux_params, allow_missing=False, force_init=True, allow_extra=False):
        if not allow_missing:
            self.init_params(initializer=None, arg_params=arg_params, aux_params=aux_params,
                             allow_missing=allow_missing, force_init=force_init, allow_extra=allow_extra)
            return
        if self.params_initialized and not force_init:
            warnings.warn("Parameters already initialized and force_init=False. set_params call ignored.", stacklevel=2)
            return
        self._exec_group.set_params(arg_params, aux_params, allow_extra=allow_extra)
        if getattr(self, "_fused", None) is not None:
            self._fused.refresh_master()
        self._params_dirty = True
        self.params_initialized = True

    def bind(self, data_shapes, label_shapes=None, for_training=True, inputs_need_grad=False, force_rebind=False,
             shared_module=None, grad_req="write"):
        if force_rebind:
            self._reset_bind()
        if self.binded:
            self.logger.warning("Already bound, ignoring bind()")
            return
        self.for_training = for_training
        self.inputs_need_grad = inputs_need_grad
        self._grad_req = grad_req
        if not for_training:
            assert not inputs_need_grad
        self._data_shapes, self._label_shapes = _parse_data_desc(self.data_names, self.label_names, data_shapes, label_shapes)
        shared_group = None
        if shared_module is not None:
            assert isinstance(shared_module, Module) and shared_module.binded and shared_module.params_initialized
            shared_group = shared_module._exec_group
            assert len(shared_group.execs) >= len(self._context)
        self._exec_group = DataParallelExecutorGroup(self._symbol, self._context, self._work_load_list, self._data_shapes,
                                                     self._label_shapes, self._param_names, for_training, inputs_need_grad,
                                                     shared_group, logger=self.logger,
                                                     fixed_param_names=self._fixed_param_names, grad_req=grad_req,
                                                     group2ctxs=self._group2ctxs, state_names=self._state_names)
        self._total_exec_bytes = 0
        self.binded = True
        if shared_module is not None:
            self.params_initialized = True
            self._arg_params = shared_module._arg_params
            self._aux_params = shared_module._aux_params
        elif self.params_initialized:
            self._exec_group.set_params(self._arg_params, self._aux_params)
        else:
            assert self._arg_params is None and self._aux_params is None
        if shared_module is not None and shared_module.optimizer_initialized:
            self.borrow_optimizer(shared_module)

    def reshape(self, data_shapes, label_shapes=None):
        assert self.binded
        self._data_shapes, self._label_shapes = _parse_data_desc(self.data_names, self.label_names, data_shapes, label_shapes)
        self._exec_group.reshape(self._data_shapes, self._label_shapes)

    def init_optimizer(self, kvstore="local", optimizer="sgd", optimizer_params=(("learning_rate", 0.01),),
                       force_init=False, initialize_from_kvstore=False, allow_missing=False, allow_extra=False):
        assert self.binded and self.params_initialized
        if self.optimizer_initialized and not force_init:
            self.logger.warning("optimizer already initialized, ignoring...")
            return
        if self._params_dirty:
            self._sync_params_from_devices()
        (kvstore, update_on_kvstore) = _create_kvstore(kvstore, len(self._context), self._arg_params)
        batch_size = self._exec_group.batch_size
        if kvstore and "dist" in kvstore.type and "_sync" in kvstore.type:
            batch_size *= kvstore.num_workers
        elif kvstore and kvstore.num_workers > 1 and "dist" not in kvstore.type:
            batch_size *= kvstore.num_workers
        rescale_grad = 1.0 / batch_size
        idx2name = {}
        if update_on_kvstore:
            idx2name.update(enumerate(self._exec_group.param_names))
        else:
            for k in range(len(self._context)):
                idx2name.update({i * len(self._context) + k: n for i, n in enumerate(self._exec_group.param_names)})
        if isinstance(optimizer, str):
            optimizer_params = dict(optimizer_params)
            if "rescale_grad" not in optimizer_params:
                optimizer_params["rescale_grad"] = rescale_grad
            optimizer = opt.create(optimizer, sym=self.symbol, param_idx2name=idx2name, **optimizer_params)
        else:
            assert isinstance(optimizer, opt.Optimizer)
            if optimizer.rescale_grad != rescale_grad:
                warnings.warn("Optimizer created manually outside Module but rescale_grad is not normalized to "
                              "1.0/batch_size/num_workers (%s vs. %s). Is this intended?" % (optimizer.rescale_grad, rescale_grad),
                              stacklevel=2)
            if not optimizer.idx2name:
                optimizer.idx2name = idx2name.copy()
        self._optimizer = optimizer
        self._kvstore = kvstore
        self._update_on_kvstore = update_on_kvstore
        self._updater = None
        if kvstore:
            if self._compression_params:
                kvstore.set_gradient_compression(self._compression_params)
            if update_on_kvstore and not initialize_from_kvstore:
                kvstore.set_optimizer(self._optimizer)
            if initialize_from_kvstore:
                # a worker joining a running job: do not init / set_optimizer, pull live params + aux states
                self._pull_from_kvstore(kvstore)
            else:
                _initialize_kvstore(kvstore=kvstore, param_arrays=self._exec_group.param_arrays,
                                    arg_params=self._arg_params, param_names=self._param_names,
                                    update_on_kvstore=update_on_kvstore,
                                    aux_arrays=self._exec_group.aux_arrays if self._aux_names else None,
                                    aux_params=self._aux_params, aux_names=self._aux_names)
        if not update_on_kvstore:
            self._updater = opt.get_updater(optimizer)
        self.optimizer_initialized = True
        self._maybe_fuse()
        if self._preload_opt_states is not None:
            self.load_optimizer_states(self._preload_opt_states)
            self._preload_opt_states = None

    # ---------------------------------------------------------------------------------- fused update path
    def _fusable(self):
        """One GPU per process, a dense optimizer the fused kernel implements, no parameter server in the loop: the
        whole ``update()`` (cross-GPU gradient sum over NVLink, rescale / clip / weight decay, optimizer, low-precision
        weight fan-out, gradient clear) is then ONE kernel instead of the reference's per-key push / pull / update ops
        (module.py:685-719 -> model.py:_update_params_on_kvstore)."""
        from ..base import getenv
        if not getenv("MXB200_FUSED_MODULE", True) or getattr(self, "_no_fuse", False):
            return False
        if len(self._context) != 1 or self._context[0].device_type != "gpu" or not torch.cuda.is_available():
            return False
        if not self.for_training or self._grad_req != "write" or self._compression_params:
            return False
        if type(self._optimizer) not in (opt.SGD, opt.NAG, opt.Adam, opt.LAMB):
            return False
        kv = self._kvstore
        if kv is not None and ("dist" in kv.type or hasattr(kv, "elastic_group")):
            return False
        if len(self._exec_group.execs) != 1:
            return False
        ex = self._exec_group.execs[0]
        for n in self._exec_group.param_names:
            a, g = ex.arg_dict.get(n), ex.grad_dict.get(n)
            if a is None or g is None or getattr(a, "stype", "default") != "default" or getattr(g, "stype", "default") != "default":
                return False
            if a._t.dtype not in (torch.float32, torch.bfloat16):      # fp16 / fp64: generic per-key updaters
                return False
        return bool(self._exec_group.param_names)

    def _maybe_fuse(self):
        if getattr(self, "_fused", None) is not None or not self._fusable():
            return
        from ..parallel.trainer import DeviceTrainer
        ex = self._exec_group.execs[0]
        ctx = self._context[0]
        kv = self._kvstore
        world = kv.num_workers if kv is not None else 1
        rank = kv.rank if kv is not None else 0
        names = list(self._exec_group.param_names)
        o = self._optimizer
        named = [(n, ex.arg_dict[n]._t.detach()) for n in names]
        # per-parameter multipliers exactly as Optimizer._get_lr / _get_wd resolve them (optimizer.py:set_lr_mult / set_wd_mult)
        lr_mult = {n: float(o.lr_mult.get(n, 1.0)) for n in names}
        wd_mult = {n: float(o.wd_mult.get(n, 1.0)) for n in names}
        dt = DeviceTrainer(named, ctx.torch_device, rank=rank, world=world, lr_mult=lr_mult, wd_mult=wd_mult)
        views = dt.params()
        for n in names:
            a = ex.arg_dict[n]
            if a._grad_hook is not None:          # arena gradients are cleared by the fused kernel itself
                a._grad_hook.remove()
                a._grad_hook = None
            a._t = views[n]
            ex.grad_dict[n]._t = dt.grad_sinks[n]
        if world > 1:
            import torch.distributed as dist
            for ar in dt.arenas:                    # rank 0's initial weights win (kv.init semantics)
                dist.broadcast(ar.weights, src=0)
            for n in self._aux_names:
                if n in ex.aux_dict:
                    dist.broadcast(ex.aux_dict[n]._t, src=0)
        self._fused = dt

    def _drop_fused(self):
        dt, self._fused = getattr(self, "_fused", None), None
        if dt is not None:
            try:
                dt.close()
            except Exception:  # noqa: BLE001
                pass

    def _fused_update(self):
        o = self._optimizer
        o.num_update += 1
        lr = o.lr_scheduler(o.num_update) if o.lr_scheduler is not None else o.lr
        kind = {opt.SGD: "sgd", opt.NAG: "nag", opt.Adam: "adam", opt.LAMB: "lamb"}[type(o)]
        kw = {}
        if kind in ("adam", "lamb"):
            kw.update(beta1=o.beta1, beta2=o.beta2, eps=o.epsilon)
        if kind == "lamb":
            kw.update(bias_correction=o.bias_correction,
                      lamb_bounds=(-1.0 if o.lower_bound is None else o.lower_bound,
                                   -1.0 if o.upper_bound is None else o.upper_bound))
        self._fused.step(kind, lr=lr, wd=o.wd, momentum=getattr(o, "momentum", 0.0), rescale=o.rescale_grad,
                         clip=o.clip_gradient, multi_precision=o.multi_precision, **kw)

    def _pull_from_kvstore(self, kvstore):
        """NEW_WORKER bootstrap (module.py:561-571): register keys without initialising them and pull values."""
        for idx, (name, arrs) in enumerate(zip(self._param_names, self._exec_group.param_arrays)):
            kvstore.init(name, self._arg_params[name])
            kvstore.pull(name, arrs, priority=-idx)
        for idx, (name, arrs) in enumerate(zip(self._aux_names, self._exec_group.aux_arrays)):
            kvstore.init(name, self._aux_params[name], True)
            kvstore.pull(name, arrs, priority=-idx)
        self._params_dirty = True

    def borrow_optimizer(self, shared_module):
        assert shared_module.optimizer_initialized
        self._optimizer = shared_module._optimizer
        self._kvstore = shared_module._kvstore
        self._update_on_kvstore = shared_module._update_on_kvstore
        self._updater = shared_module._updater
        self.optimizer_initialized = True

    def forward(self, data_batch, is_train=None):
        assert self.binded and self.params_initialized
        curr_data_shapes = tuple(i.shape for i in self._data_shapes)
        if isinstance(data_batch, list):
            new_data_shapes = tuple(b.data[0].shape for b in data_batch)
        else:
            new_data_shapes = tuple(i.shape for i in data_batch.data)
        if curr_data_shapes != new_data_shapes:
            if hasattr(data_batch, "provide_data") and data_batch.provide_data:
                new_dshape = data_batch.provide_data
            else:
                new_dshape = [DataDesc(i.name, shape, i.dtype, i.layout) for i, shape in zip(self._data_shapes, new_data_shapes)]
            if hasattr(data_batch, "provide_label") and data_batch.provide_label:
                new_lshape = data_batch.provide_label
            elif hasattr(data_batch, "label") and data_batch.label:
                new_lshape = [DataDesc(i.name, j.shape, i.dtype, i.layout) for i, j in zip(self._label_shapes, data_batch.label)]
            else:
                new_lshape = None
            self.reshape(new_dshape, new_lshape)
        self._exec_group.forward(data_batch, is_train)

    def backward(self, out_grads=None):
        assert self.binded and self.params_initialized
        self._exec_group.backward(out_grads=out_grads)

    def update(self, update_aux_params=False):
        """Apply the optimizer to the gradients of the last forward-backward (module.py:685-719)."""
        assert self.binded and self.params_initialized and self.optimizer_initialized
        self._params_dirty = True
        if getattr(self, "_fused", None) is not None:
            self._fused_update()
            if update_aux_params:
                self.store_aux_params()
            return
        if self._update_on_kvstore:
            _update_params_on_kvstore(self._exec_group.param_arrays, self._exec_group.grad_arrays, self._kvstore,
                                      self._exec_group.param_names)
        else:
            _update_params(self._exec_group.param_arrays, self._exec_group.grad_arrays, updater=self._updater,
                           num_device=len(self._context), kvstore=self._kvstore,
                           param_names=self._exec_group.param_names)
        if update_aux_params:
            self.store_aux_params()

    def store_aux_params(self):
        """Push aux states to the store (averaged across workers via exclude-update keys) and pull them back."""
        if self._kvstore is None or not self._aux_names:
            return
        _store_aux_params_on_kvstore(self._exec_group.aux_arrays, self._kvstore, self._aux_names)

    def get_outputs(self, merge_multi_context=True):
        assert self.binded and self.params_initialized
        return self._exec_group.get_outputs(merge_multi_context=merge_multi_context)

    def get_input_grads(self, merge_multi_context=True):
        assert self.binded and self.params_initialized and self.inputs_need_grad
        return self._exec_group.get_input_grads(merge_multi_context=merge_multi_context)

    def get_states(self, merge_multi_context=True):
        assert self.binded and self.params_initialized
        return self._exec_group.get_states(merge_multi_context=merge_multi_context)

    def set_states(self, states=None, value=None):
        assert self.binded and self.params_initialized
        self._exec_group.set_states(states, value)

    def update_metric(self, eval_metric, labels, pre_sliced=False):
        self._exec_group.update_metric(eval_metric, labels, pre_sliced)

    def _sync_params_from_devices(self):
        self._exec_group.get_params(self._arg_params, self._aux_params)
        if self._kvstore and self._update_on_kvstore:
            for param_name, param_val in sorted(self._arg_params.items()):
                if getattr(param_val, "stype", "default") == "row_sparse":
                    row_ids = nd.arange(0, param_val.shape[0], dtype="int64")
                    self._kvstore.row_sparse_pull(param_name, param_val, row_ids=row_ids)
        self._params_dirty = False

    def save_optimizer_states(self, fname):
        assert self.optimizer_initialized
        if getattr(self, "_fused", None) is not None:
            import pickle
            dt = self._fused
            st = dt.export_state()                       # collective: un-shards the ZeRO-1 slices over peer memory
            blob = {"num_update": self._optimizer.num_update,
                    "params": {n: {k: v.cpu() for k, v in d.items() if k != "weight"} for n, d in st["params"].items()}}
            with open(fname, "wb") as fout:
                pickle.dump({"mxb200_fused_states": blob}, fout)
            return
        if self._update_on_kvstore:
            self._kvstore.save_optimizer_states(fname)
        else:
            with open(fname, "wb") as fout:
                fout.write(self._updater.get_states())

    def load_optimizer_states(self, fname):
        assert self.optimizer_initialized
        if getattr(self, "_fused", None) is not None:
            import pickle
            with open(fname, "rb") as f:
                blob = pickle.load(f)
            if not (isinstance(blob, dict) and "mxb200_fused_states" in blob):
                raise ValueError("%s was not written by the fused update path; set MXB200_FUSED_MODULE=0 to load it" % fname)
            blob = blob["mxb200_fused_states"]
            dt = self._fused
            dev = dt.device
            cur = dt.export_state()["params"]
            state = {"num_update": blob["num_update"],
                     "params": {n: dict({"weight": cur[n]["weight"]}, **{k: v.to(dev) for k, v in d.items()})
                                for n, d in blob["params"].items() if n in cur}}
            dt.import_state(state)
            self._optimizer.num_update = int(blob["num_update"])
            return
        if self._update_on_kvstore:
            self._kvstore.load_optimizer_states(fname)
        else:
            with open(fname, "rb") as f:
                self._updater.set_states(f.read())

    def install_monitor(self, mon):
        assert self.binded
        self._exec_group.install_monitor(mon)

    def prepare(self, data_batch, sparse_row_id_fn=None):
        assert self.binded
        if sparse_row_id_fn is not None:
            if not self._kvstore or not self._update_on_kvstore:
                warnings.warn(UserWarning("Parameters are not updated in the KVStore. No need to call sparse_row_id_fn."))
            else:
                row_ids = sparse_row_id_fn(data_batch)
                assert isinstance(row_ids, dict), "Expected dict output from sparse_row_id_fn"
                for param_name, row_id in row_ids.items():
                    param_idx = self._exec_group.param_names.index(param_name)
                    param_val = self._exec_group.param_arrays[param_idx]
                    self._kvstore.row_sparse_pull(param_name, param_val, row_ids=[row_id] * len(param_val), priority=-param_idx)
