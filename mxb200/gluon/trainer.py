"""gluon.Trainer (reference: python/mxnet/gluon/trainer.py:62-423).

Two execution paths behind one API:

* **fused** (default on CUDA for SGD / NAG / Adam / LAMB with dense parameters on one GPU per process):
  parameters and gradients are re-homed into symmetric arenas and ``step`` is ONE kernel launch per dtype
  group that performs the cross-GPU gradient reduction (P2P over NVLink), rescale/clip, the optimizer
  update with fp32 master weights and the weight fan-out (mxb200/parallel/trainer.py,
  csrc/cuda/allreduce.cu).  This replaces _allreduce_grads + _update (trainer.py:304-374).
* **generic**: the reference algorithm -- kvstore push/pull per parameter + per-context Updaters -- used on
  CPU, with several contexts in one process, with sparse gradients or custom optimizers.

Provenance: adapted from Apache MXNet (incubating) 1.3.1 ``python/mxnet/gluon/trainer.py`` (Apache License 2.0).  The public API, argument
names and control flow follow the original so that user scripts keep working; the tensor maths runs on torch /
the mxb200 kernels.  This file is API-compatibility glue, not original B200 work ("Provenance").
"""
import warnings

import torch

from .. import optimizer as opt
from ..base import getenv
from ..model import _create_kvstore
from ..ndarray import NDArray, sparse as _sp
from .parameter import ParameterDict, Parameter

__all__ = ["Trainer"]


class Trainer(object):
    def __init__(self, params, optimizer, optimizer_params=None, kvstore="device", compression_params=None,
                 update_on_kvstore=None):
        if isinstance(params, (dict, ParameterDict)):
            params = list(params.values())
        if not isinstance(params, (list, tuple)):
            raise ValueError("First argument must be a list or dict of Parameters, got %s." % (type(params)))
        self._params = []
        self._contains_sparse = False
        self._param2idx = {}
        for i, param in enumerate(params):
            if not isinstance(param, Parameter):
                raise ValueError("First argument must be a list or dict of Parameters, got list of %s." % (type(param)))
            self._param2idx[param.name] = i
            self._params.append(param)
            param._set_trainer(self)
            if param._stype != "default" or param._grad_stype != "default":
                self._contains_sparse = True
        self._compression_params = compression_params
        optimizer_params = optimizer_params if optimizer_params else {}
        self._scale = float(optimizer_params.get("rescale_grad", 1.0))
        self._contexts = self._check_contexts()
        self._init_optimizer(optimizer, optimizer_params)
        self._kvstore_params = {"kvstore": kvstore, "update_on_kvstore": update_on_kvstore}
        self._kv_initialized = False
        self._kvstore = None
        self._update_on_kvstore = None
        self._params_to_init = []
        self._fused = None
        self._reset_kvstore()

    def _check_contexts(self):
        contexts = None
        for param in self._params:
            ctx = param.list_ctx()
            assert contexts is None or contexts == ctx, \
                "All Parameters must be initialized on the same set of contexts, but Parameter %s is initialized on " \
                "%s while previous Parameters are initialized on %s." % (param.name, str(ctx), str(contexts))
            contexts = ctx
        return contexts

    def _init_optimizer(self, optimizer, optimizer_params):
        param_dict = {i: param for i, param in enumerate(self._params)}
        if isinstance(optimizer, opt.Optimizer):
            assert not optimizer_params, "optimizer_params must be None if optimizer is an instance of Optimizer " \
                                         "instead of str"
            self._optimizer = optimizer
            self._optimizer.param_dict = param_dict
        else:
            self._optimizer = opt.create(optimizer, param_dict=param_dict, **optimizer_params)
        self._updaters = [opt.get_updater(self._optimizer) for _ in self._contexts]

    def _reset_kvstore(self):
        if self._kvstore and "dist" in self._kvstore.type:
            raise RuntimeError("Cannot reset distributed KVStore.")
        self._kv_initialized = False
        self._kvstore = None
        self._update_on_kvstore = None
        self._params_to_init = [param for param in self._params]
        if self._fused is not None:
            self._unfuse()

    # ---------------------------------------------------------------------------------- kvstore init
    def _init_kvstore(self):
        config = self._kvstore_params
        sparse_weight = any(p._stype != "default" for p in self._params)
        sparse_grad = any(p._grad_stype != "default" for p in self._params)
        if sparse_weight:
            # row_sparse weights live in the kvstore and are updated there, even with one device
            # (gluon/trainer.py:168-180 of the reference: _create_sparse_kvstore)
            from .. import kvstore as _kvs
            kv = config["kvstore"]
            kvstore = _kvs.create(kv) if isinstance(kv, str) else kv
            if kvstore is None:
                raise ValueError("Cannot create a kvstore for sparse weights from %r" % (kv,))
            if config["update_on_kvstore"] is False:
                raise RuntimeError("Cannot set update_on_kvstore to False when sparse weights are present.")
            update_on_kvstore = True
        elif sparse_grad:
            # dense weights with row_sparse gradients: aggregated through the kvstore, updated locally (reference :181-185)
            kvstore, _ = _create_kvstore(config["kvstore"], len(self._contexts), {})
            update_on_kvstore = False
            if config["update_on_kvstore"]:
                update_on_kvstore = True
        else:
            arg_arrays = {}
            for param in self._params:
                try:
                    arg_arrays[param.name] = param._raw_data(self._contexts[0])
                except Exception:  # noqa: BLE001
                    pass
            kvstore, update_on_kvstore = _create_kvstore(config["kvstore"], len(self._contexts), arg_arrays)
            if config["update_on_kvstore"] is not None:
                update_on_kvstore = config["update_on_kvstore"]
        if kvstore:
            if self._compression_params:
                kvstore.set_gradient_compression(self._compression_params)
            if "dist" in kvstore.type:
                update_on_kvstore = True if ("async" in kvstore.type or sparse_weight or sparse_grad) else update_on_kvstore
            if update_on_kvstore:
                kvstore.set_optimizer(self._optimizer)
            self._kvstore = kvstore
            self._update_on_kvstore = update_on_kvstore
        else:
            self._kvstore = None
            self._update_on_kvstore = None
        self._kv_initialized = True
        self._maybe_fuse()

    # ---------------------------------------------------------------------------------- fused path
    def _fusable(self):
        if not getenv("MXB200_FUSED_TRAINER", True):
            return False
        if self._contains_sparse or self._compression_params or len(self._contexts) != 1:
            return False
        if self._contexts[0].device_type != "gpu" or not torch.cuda.is_available():
            return False
        if type(self._optimizer) not in (opt.SGD, opt.NAG, opt.Adam, opt.LAMB):
            return False
        kv = self._kvstore
        if kv is not None and "dist" in kv.type:
            return False
        for p in self._params:
            if p._data is None:
                return False
            # the fused kernel is instantiated for fp32 and bf16 arenas; float16 / float64 nets take the generic updaters
            if p.grad_req != "null" and p._data[0]._t.dtype not in (torch.float32, torch.bfloat16):
                return False
        return True

    def init_kvstore(self):
        """Create the kvstore (and the fused arena path) now instead of lazily at the first ``step``.  Elastic jobs
        call this before the first forward pass so that a newcomer starts from the live weights."""
        if not self._kv_initialized:
            self._init_kvstore()

    def _maybe_fuse(self, state=None):
        if self._fused is not None or not self._fusable():
            return
        from ..parallel.trainer import DeviceTrainer
        ctx = self._contexts[0]
        kv = self._kvstore
        world = kv.num_workers if kv is not None else 1
        rank = kv.rank if kv is not None else 0
        elastic = kv is not None and hasattr(kv, "elastic_group")
        train = [p for p in self._params if p.grad_req != "null"]
        if not train:
            return
        named = [(p.name, p._raw_data(ctx)._t.detach()) for p in train]
        lr_mult = {p.name: p.lr_mult for p in train}
        wd_mult = {p.name: p.wd_mult for p in train}
        dt = DeviceTrainer(named, ctx.torch_device, rank=rank, world=world, lr_mult=lr_mult, wd_mult=wd_mult,
                           group=kv.peer_group(ctx.torch_device) if elastic else None)
        views = dt.params()
        stale = elastic and (state is not None or kv.is_new_worker)   # gradients computed before the weights were synced
        for p in train:
            v = views[p.name]
            d = p._data[0]
            if d._grad_hook is not None:      # arena grads are cleared by the fused kernel itself
                d._grad_hook.remove()
                d._grad_hook = None
            sink = dt.grad_sinks[p.name]
            with torch.no_grad():
                if not stale and p._grad is not None and p._grad[0]._t is not None:
                    sink.copy_(p._grad[0]._t)           # gradients of the step that triggered the lazy fusion
            d._t = v
            p._grad[0]._t = sink                        # fp32 arena view (what ``param.grad()`` shows)
            p._grad_sink = dt.grad_sinks[p.name]
        if elastic:
            if state is not None:
                dt.import_state(state, keep_full=True)
            dt.sync_from_rank0(kv._barrier)              # weights + optimizer state of rank 0 (a newcomer's live state)
            dt.broadcast_from_rank0([p._data[0]._t for p in self._params if p.grad_req == "null" and p._data is not None],
                                    kv._barrier)
            # the update counter drives lr schedules and Adam / LAMB bias corrections: newcomers continue at the job's count
            cnt = torch.tensor([float(self._optimizer.num_update), float(dt.num_update)], device=dt.device)
            dt.broadcast_from_rank0([cnt])
            self._optimizer.num_update = int(round(cnt[0].item()))
            dt.num_update = int(round(cnt[1].item()))
            kv._register_trainer(self)
        elif world > 1:
            # non-trainable state (BN running stats): start from rank 0's values (kv.init semantics)
            import torch.distributed as dist
            for p in self._params:
                if p.grad_req == "null":
                    dist.broadcast(p._data[0]._t, src=0)
            # weights: rank 0 wins
            for ar in dt.arenas:
                dist.broadcast(ar.weights, src=0)
        self._fused = dt
        self._params_to_init = []

    # elastic membership change (kvstore('elastic_device')): called by the store around the scheduler barrier
    def _elastic_detach(self):
        self._elastic_state = None
        if self._fused is None:
            return
        kv = self._kvstore
        # BatchNorm running statistics: average over the current members (one launch of the fused kernel), so that
        # newcomers -- and everybody after the change -- continue from the job-wide statistics
        aux = [p._data[0]._t for p in self._params if p.grad_req == "null" and p._data is not None
               and p._data[0]._t.is_floating_point()]
        self._fused.average_across_ranks(aux)
        # un-shard the ZeRO-1 optimizer state while everyone is still here (collective over peer memory)
        self._elastic_state = self._fused.export_state(kv._barrier)
        torch.cuda.synchronize(self._fused.device)
        kv._barrier()
        self._unfuse()

    def _elastic_attach(self):
        state, self._elastic_state = getattr(self, "_elastic_state", None), None
        self._maybe_fuse(state=state if state is not None else {"num_update": 0, "params": {}})

    def _on_membership_change(self, old_world, new_world):
        """Re-apply the dynamic-mini-batch learning-rate rule after the worker set changed (hook location in the
        reference: python/mxnet/module/base_module.py:540-549; rule: lr_scheduler.DynamicMiniBatchScheduler).  With a
        constant per-worker batch the global batch follows the membership; an abrupt k-fold lr change also rescales the
        accumulated momentum (momentum correction)."""
        o = self._optimizer
        sched = getattr(o, "lr_scheduler", None)
        if sched is None or not hasattr(sched, "on_membership_change") or old_world == new_world:
            return None
        per_worker = getattr(sched, "per_worker_batch", None)
        new_batch = per_worker * new_world if per_worker else sched.base_batch
        lr_before = sched(o.num_update) if o.num_update else sched.base_lr * sched.cur_factor
        sched.on_membership_change(new_batch, num_update=o.num_update)
        corr = sched.momentum_correction() if sched.smooth_steps == 0 else 1.0
        est = getattr(self, "_elastic_state", None)
        if est is not None and corr != 1.0 and getattr(o, "momentum", 0.0):
            # between detach and attach of an elastic re-init: correct the exported (full) state, so that every rank --
            # newcomers included, replicated one-shot arenas included -- continues from the SAME corrected momentum
            for d in est["params"].values():
                if "mom" in d:
                    d["mom"].mul_(corr)
        elif self._fused is not None and corr != 1.0 and getattr(o, "momentum", 0.0):
            self._fused.scale_state("mom", corr)
        elif corr != 1.0:
            for upd in self._updaters:
                for st in upd.states.values():
                    for x in (st if isinstance(st, (list, tuple)) else [st]):
                        if isinstance(x, NDArray) and x._t.dtype.is_floating_point and getattr(o, "momentum", 0.0):
                            x._t.mul_(corr)
        self.last_lr_change = {"old_world": old_world, "new_world": new_world, "global_batch": new_batch,
                               "lr_before": lr_before, "lr_after": sched(max(1, o.num_update)),
                               "momentum_correction": corr}
        return self.last_lr_change

    def _unfuse(self):
        dt, self._fused = self._fused, None
        for p in self._params:
            p._grad_sink = None
            if p.grad_req == "null" or p._data is None:
                continue
            d = p._data[0]
            t = d._t.detach().clone().contiguous() if d._t.is_contiguous() else d._t.detach().clone()
            d._t = t
            d._grad_req = None
            d.attach_grad(p.grad_req)
            p._grad[0]._t = d._t.grad
        try:
            dt.close()
        except Exception:  # noqa: BLE001
            pass

    def _fused_barrier(self):
        kv = self._kvstore
        if kv is not None and hasattr(kv, "_barrier") and hasattr(kv, "elastic_group"):
            return kv._barrier
        if self._fused is not None and self._fused.world > 1:
            import torch.distributed as dist
            return dist.barrier
        return None

    def check_health(self):
        """Raise if a fused collective timed out (a peer never arrived)."""
        if self._fused is not None:
            torch.cuda.synchronize(self._fused.device)
            self._fused.check()

    # ---------------------------------------------------------------------------------- properties
    @property
    def learning_rate(self):
        if not isinstance(self._optimizer, opt.Optimizer):
            raise UserWarning("Optimizer has to be defined before its learning rate can be accessed.")
        return self._optimizer.learning_rate

    def set_learning_rate(self, lr):
        if not isinstance(self._optimizer, opt.Optimizer):
            raise UserWarning("Optimizer has to be defined before its learning rate is mutated.")
        self._optimizer.set_learning_rate(lr)

    def _init_params(self):
        assert self._kv_initialized, "Cannot initialize parameters in KVStore when KVStore is not initialized."
        params_to_init = []
        if self._kvstore:
            for param in self._params_to_init:
                if param._deferred_init:
                    params_to_init.append(param)
                else:
                    param_arrays = param._check_and_get(param._data, list)
                    idx = self._param2idx[param.name]
                    self._kvstore.init(idx, param_arrays[0])
                    if param._stype == "default":
                        self._kvstore.pull(idx, param_arrays, priority=-idx)
        self._params_to_init = params_to_init

    def _row_sparse_pull(self, parameter, out, row_id):
        if not self._kv_initialized:
            self._init_kvstore()
        if self._params_to_init:
            self._init_params()
        idx = self._param2idx[parameter.name]
        self._kvstore.row_sparse_pull(idx, out=out, row_ids=row_id, priority=-idx)

    # ---------------------------------------------------------------------------------- step
    def step(self, batch_size, ignore_stale_grad=False):
        """One optimisation step; gradients are normalised by 1/batch_size (trainer.py:254-280)."""
        rescale_grad = self._scale / batch_size
        if not self._kv_initialized:
            self._init_kvstore()
        if self._fused is not None:
            self._fused_step(rescale_grad)
            return
        if self._params_to_init:
            self._init_params()
        self._optimizer.rescale_grad = rescale_grad
        self._allreduce_grads()
        self._update(ignore_stale_grad)

    def _fused_step(self, rescale_grad):
        o = self._optimizer
        o.rescale_grad = rescale_grad
        phase = getattr(self, "_graph_phase", None)      # None | "host" | "device" (parallel/graph.py)
        if phase != "device":
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
        self._fused.step(kind, lr=lr, wd=o.wd, momentum=getattr(o, "momentum", 0.0), rescale=rescale_grad,
                         clip=o.clip_gradient, multi_precision=o.multi_precision, phase=phase, **kw)

    def allreduce_grads(self):
        if not self._kv_initialized:
            self._init_kvstore()
        if self._fused is not None:
            raise RuntimeError("allreduce_grads()/update() split is not available on the fused path; construct the "
                               "Trainer with update_on_kvstore=False and set MXB200_FUSED_TRAINER=0, or call step().")
        if self._params_to_init:
            self._init_params()
        assert not (self._kvstore and self._update_on_kvstore), \
            "allreduce_grads() when parameters are updated on kvstore is not supported. Try setting " \
            "`update_on_kvstore` to False when creating trainer."
        self._allreduce_grads()

    def _allreduce_grads(self):
        if self._kvstore:
            live = [(i, param) for i, param in enumerate(self._params) if param.grad_req != "null"]
            for i, param in live:
                self._kvstore.push(i, param._raw_list_grad(), priority=-i)
            if not self._update_on_kvstore:
                # all pushes first: a multi-GPU device store then reduces every key in ONE bucketed fused launch per GPU
                for i, param in live:
                    self._kvstore.pull(i, param._raw_list_grad(), priority=-i, ignore_sparse=False)

    def update(self, batch_size, ignore_stale_grad=False):
        if not self._kv_initialized:
            self._init_kvstore()
        if self._params_to_init:
            self._init_params()
        assert not (self._kvstore and self._update_on_kvstore), \
            "update() when parameters are updated on kvstore is not supported. Try setting `update_on_kvstore` to " \
            "False when creating trainer."
        self._optimizer.rescale_grad = self._scale / batch_size
        self._update(ignore_stale_grad)

    def _update(self, ignore_stale_grad=False):
        for i, param in enumerate(self._params):
            if param.grad_req == "null":
                continue
            if self._kvstore and self._update_on_kvstore:
                # row_sparse parameters: the reference pulls the rows lazily from row_sparse_data(); the dense backing
                # arrays kept here are refreshed in full instead (single box: a device copy)
                self._kvstore.pull(i, param._raw_list_data(), priority=-i, ignore_sparse=False)
                continue
            for upd, arr, grad in zip(self._updaters, param._raw_list_data(), param.list_grad()):
                upd(i, grad, arr)

    # ---------------------------------------------------------------------------------- states
    def save_states(self, fname):
        assert self._optimizer is not None
        if not self._kv_initialized:
            self._init_kvstore()
        if self._fused is not None:
            # collective on the fused multi-GPU path: every rank's ZeRO-1 slice is fetched over peer memory, then each
            # rank could write the same complete file (callers usually let rank 0 write).  Layout: per-parameter tensors,
            # independent of arena packing and world size; a pickle like the reference's updater states.
            import pickle
            st = self._fused.export_state(self._fused_barrier())
            blob = {"num_update": self._optimizer.num_update, "trainer_num_update": self._fused.num_update,
                    "params": {n: {k: v.cpu() for k, v in d.items() if k != "weight"} for n, d in st["params"].items()}}
            with open(fname, "wb") as fout:
                pickle.dump({"mxb200_fused_states": blob}, fout)
            return
        if self._params_to_init:
            self._init_params()
        if self._update_on_kvstore:
            self._kvstore.save_optimizer_states(fname, dump_optimizer=True)
        else:
            with open(fname, "wb") as fout:
                fout.write(self._updaters[0].get_states(dump_optimizer=True))

    def load_states(self, fname):
        if not self._kv_initialized:
            self._init_kvstore()
        if self._fused is not None:
            import pickle
            with open(fname, "rb") as f:
                raw = f.read()
            try:
                blob = pickle.loads(raw)
            except Exception:  # noqa: BLE001
                blob = None
            if isinstance(blob, dict) and "mxb200_fused_states" in blob:
                blob = blob["mxb200_fused_states"]
                dt = self._fused
                dev = dt.device
                cur = {n: {"weight": v.detach()} for n, v in dt.params().items()}
                saved = blob["params"]
                if not any(n in cur for n in saved) and len(saved) == len(cur):
                    # other name prefixes (a re-created net): match by position, like the reference's index-keyed states
                    saved = {n: d for n, d in zip(cur.keys(), saved.values())}
                state = {"num_update": int(blob.get("trainer_num_update", blob["num_update"])),
                         "params": {n: dict(cur[n], **{k: v.to(dev) for k, v in d.items()})
                                    for n, d in saved.items() if n in cur}}
                dt.import_state(state)
                self._optimizer.num_update = int(blob["num_update"])
                return
            # a file written by the generic (per-parameter updater) path, e.g. by the reference: adopt it there
            self._unfuse()
            self._fusable = lambda: False
        if self._params_to_init:
            self._init_params()
        if self._update_on_kvstore:
            self._kvstore.load_optimizer_states(fname)
            self._optimizer = self._kvstore._updater.optimizer
        else:
            with open(fname, "rb") as f:
                states = f.read()
            for updater in self._updaters:
                updater.set_states(states)
                updater.optimizer = self._updaters[0].optimizer
            self._optimizer = self._updaters[0].optimizer
        param_dict = {i: param for i, param in enumerate(self._params)}
        self._optimizer.param_dict = param_dict
