"""KVStore: key-value store for data-parallel training (reference API: python/mxnet/kvstore.py:116-688,
C++ include/mxnet/kvstore.h:60-446, src/kvstore/kvstore_local.h, comm.h).

Semantics preserved: ``push`` = reduce the per-device values of a key (duplicates of a key
in one call are summed) then either run the updater on the stored value or replace it; ``pull`` =
broadcast the stored value into every output; ``init`` once per key; int or str keys but never mixed; keys
initialised with ``exclude_update=True`` are numbered from ``MAX_ALLOWED_KEY_FOR_UPDATE+1`` and are
*averaged* instead of optimised (the elastic fork's aux-state keys, kvstore_local.h:100-120,
kvstore_dist_server.h:345-379).

Store types (substring matched like src/kvstore/kvstore.cc:40-77):
  local / local_update_cpu / local_allreduce_cpu : host reduce  (CommCPU, comm.h:103-441)
  device / local_allreduce_device / nccl          : device reduce.  Inside one process the per-device values
        are summed on the owner device with our ElementwiseSum kernel after P2P copies.  When the job runs one
        process per GPU (torchrun / tools/launch.py) the store spans the processes: rank/num_workers come from
        the process group and gluon.Trainer / Module route whole steps through the fused
        allreduce+update kernel (mxb200/parallel/trainer.py) instead of per-key push/pull.
  dist_sync / dist_device_sync / dist_async       : parameter-server path (kvstore/dist.py)
"""
import pickle

import torch

from ..base import MXNetError, getenv
from ..ndarray import NDArray, sparse as _sp
from ..ndarray.ndarray import zeros_like
from .. import optimizer as opt
from .compression import GradientCompression

__all__ = ["KVStore", "KVStoreLocal", "create", "MAX_ALLOWED_KEY_FOR_UPDATE"]

MAX_ALLOWED_KEY_FOR_UPDATE = 10000000


def _as_list(x):
    return list(x) if isinstance(x, (list, tuple)) else [x]


def _key_value_lists(key, value):
    """Normalise (key, value) into parallel lists: keys[i] <-> list of per-device arrays."""
    if isinstance(key, (int, str)):
        vals = _as_list(value)
        return [key], [vals]
    keys = list(key)
    if len(keys) == 0:
        return [], []
    assert isinstance(value, (list, tuple)), "a list of keys needs a list of values"
    if len(value) == len(keys) and not isinstance(value[0], (list, tuple)):
        return keys, [[v] for v in value]
    if len(value) == len(keys):
        return keys, [list(v) for v in value]
    # flat list: k values per key
    assert len(value) % len(keys) == 0, "the number of values must be a multiple of the number of keys"
    n = len(value) // len(keys)
    return keys, [list(value[i * n:(i + 1) * n]) for i in range(len(keys))]


class KVStore(object):
    """Abstract interface (python/mxnet/kvstore.py).  Concrete: KVStoreLocal, KVStoreDist."""

    def __init__(self, kv_type):
        self._type = kv_type
        self._updater = None
        self._updater_func = None
        self._str2int = {}
        self._key_type = None
        self._next_str_key = [0, MAX_ALLOWED_KEY_FOR_UPDATE + 1]
        self._compression = None
        self._barrier_before_exit = True

    # identity -------------------------------------------------------------------------------------
    @property
    def type(self):
        return self._type

    @property
    def rank(self):
        return 0

    @property
    def num_workers(self):
        return 1

    # key typing -----------------------------------------------------------------------------------
    def _set_key_type(self, key):
        t = str if isinstance(key, str) else int
        if self._key_type is None:
            self._key_type = t
        elif self._key_type is not t:
            raise MXNetError("inconsistent types of keys detected: a store uses either int or str keys, not both")

    def _int_key(self, key, create=False, exclude_update=False):
        """str keys get dense integer ids; exclude-update (aux) keys live above MAX_ALLOWED_KEY_FOR_UPDATE."""
        self._set_key_type(key)
        if isinstance(key, int):
            return key
        if key not in self._str2int:
            if not create:
                raise MXNetError("key %s has not been initialized" % key)
            slot = 1 if exclude_update else 0
            self._str2int[key] = self._next_str_key[slot]
            self._next_str_key[slot] += 1
        return self._str2int[key]

    # API ------------------------------------------------------------------------------------------
    def init(self, key, value, exclude_update=False):
        raise NotImplementedError

    def push(self, key, value, priority=0):
        raise NotImplementedError

    def pull(self, key, out=None, priority=0, ignore_sparse=True):
        raise NotImplementedError

    def row_sparse_pull(self, key, out=None, priority=0, row_ids=None):
        raise NotImplementedError

    def set_gradient_compression(self, compression_params):
        if ("device" not in self._type) and ("dist" not in self._type):
            raise MXNetError("Gradient compression is not supported for this type of kvstore")
        self._compression = GradientCompression(compression_params)

    def set_optimizer(self, optimizer):
        self._set_updater(opt.get_updater(optimizer))

    def _set_updater(self, updater):
        self._updater = updater
        self._updater_func = updater

    def save_optimizer_states(self, fname, dump_optimizer=False):
        assert self._updater is not None, "Cannot save states for distributed training"
        with open(fname, "wb") as fout:
            fout.write(self._updater.get_states(dump_optimizer))

    def load_optimizer_states(self, fname):
        assert self._updater is not None, "Cannot load states for distributed training"
        with open(fname, "rb") as f:
            self._updater.set_states(f.read())

    def _barrier(self):
        pass

    def _membership_change_barrier(self, env):
        raise MXNetError("membership change barrier is only available for dist / multi-process kvstores "
                         "(python/mxnet/kvstore.py:617-624)")

    def _send_command_to_servers(self, head, body):
        pass

    def get_num_dead_node(self, node_id=0, timeout=60):
        return 0


class KVStoreLocal(KVStore):
    """Single-process store: values of one key are one array per device (KVStoreLocal + CommCPU/CommDevice)."""

    def __init__(self, kv_type="local"):
        super(KVStoreLocal, self).__init__(kv_type)
        self._store = {}
        self._device_mode = ("device" in kv_type) or ("nccl" in kv_type)
        self._residuals = {}
        self._fused = {}            # device set -> FusedLocalComm (bucketed push / pull on the fused kernel)
        self._plans = {}            # (key, devices) -> cached routing decision of the hot per-key calls
        self._pull_q = None         # deferred pulls of fused keys: per device ([dst], [src]) drained by one foreach copy

    # reduce / broadcast (Comm) ----------------------------------------------------------------------
    def _reduce(self, ikey, vals):
        """Sum the per-device values.  Returns an array on the 'owner' device (CommDevice) or pinned/host
        memory (CommCPU)."""
        if any(isinstance(v, _sp.RowSparseNDArray) for v in vals):
            tgt = self._store[ikey].context if ikey in self._store else vals[0].context
            return _sp.add_n_rsp([v.as_in_context(tgt) for v in vals])
        if len(vals) == 1:
            return vals[0]
        if self._compression is not None and self._device_mode:
            return self._reduce_compressed(ikey, vals)
        if self._device_mode and vals[0]._t.is_cuda and getenv("MXNET_KVSTORE_USETREE", False) and \
                len({v._t.device for v in vals}) == len(vals) > 2:
            # topology-aware tree reduce (CommDeviceTree, comm_tree.h; trees from kvstore/gpu_topology.py)
            devs = tuple(v._t.device.index for v in vals)
            tc = self._tree_comms.get(devs) if hasattr(self, "_tree_comms") else None
            if tc is None:
                from .gpu_topology import TreeComm
                if not hasattr(self, "_tree_comms"):
                    self._tree_comms = {}
                tc = self._tree_comms[devs] = TreeComm(devs)
            with torch.no_grad():
                root = ikey % len(vals) if isinstance(ikey, int) else 0     # spread the roots over the devices
                return NDArray(tc.reduce([v._t for v in vals], root))
        if self._device_mode and vals[0]._t.is_cuda:
            owner = self._store[ikey]._t.device if ikey in self._store and self._store[ikey]._t.is_cuda else vals[0]._t.device
            with torch.no_grad():
                ts = [v._t if v._t.device == owner else v._t.to(owner, non_blocking=True) for v in vals]
                out = torch.empty_like(ts[0])
                if all(t.is_contiguous() for t in ts):
                    from ..ops import optim as KO
                    with torch.cuda.device(owner):
                        KO.elementwise_sum(out, ts)
                else:
                    out = torch.stack(ts).sum(0)
            return NDArray(out)
        # host reduce
        with torch.no_grad():
            acc = vals[0]._t.detach().to("cpu", copy=True).float()
            for v in vals[1:]:
                acc += v._t.detach().to("cpu").float()
        return NDArray(acc.to(vals[0]._t.dtype), vals[0].context if not vals[0]._t.is_cuda else None)

    def _reduce_compressed(self, ikey, vals):
        """2-bit quantise on the source device, move the small buffer, dequantise and sum at the owner
        (CommDevice::ReduceCompressed, comm.h:552-596)."""
        owner = vals[0]._t.device
        acc = None
        for i, v in enumerate(vals):
            res = self._residuals.setdefault((ikey, i), torch.zeros_like(v._t, dtype=torch.float32))
            q = self._compression.quantize(v._t, res)
            d = self._compression.dequantize(q.to(owner), v._t.numel(), v._t.shape, owner)
            acc = d if acc is None else acc + d
        return NDArray(acc.to(vals[0]._t.dtype), vals[0].context)

    # API --------------------------------------------------------------------------------------------
    def init(self, key, value, exclude_update=False):
        keys, vals = _key_value_lists(key, value)
        for k, vlist in zip(keys, vals):
            ik = self._int_key(k, create=True, exclude_update=exclude_update)
            if ik in self._store:
                raise MXNetError("duplicate init of key %s" % str(k))
            v = vlist[0]
            self._store[ik] = v.copy() if isinstance(v, (NDArray, _sp.BaseSparseNDArray)) else v

    def _group(self, keys, vals):
        """GroupKVPairs: duplicates of a key in one call are merged (kvstore_local.h:349-379)."""
        grouped, order = {}, []
        for k, vlist in zip(keys, vals):
            ik = self._int_key(k)
            if ik not in grouped:
                grouped[ik] = []
                order.append(ik)
            grouped[ik].extend(vlist)
        return [(ik, grouped[ik]) for ik in order]

    # -- fused multi-GPU path (one process, one array per GPU): kvstore/fused_local.py ---------------------
    def _fused_comm(self, vlist):
        """The bucketed fused-kernel communicator for this device set, or None when the request does not qualify."""
        if not self._device_mode or self._compression is not None:
            return None
        from .fused_local import FusedLocalComm
        devs = FusedLocalComm.usable(vlist)
        if devs is None:
            return None
        key = tuple(sorted(d.index for d in devs))
        fc = self._fused.get(key)
        if fc is None:
            try:
                fc = FusedLocalComm([torch.device("cuda", i) for i in key])
            except Exception:  # noqa: BLE001 -- no peer access between these GPUs: per-key path
                fc = False
            self._fused[key] = fc
        return fc or None

    def _flush_fused(self):
        for fc in self._fused.values():
            if fc and fc.pending:
                iks = list(fc.pending)
                fc.flush(self._store)
                for ik in iks:
                    stored = self._store[ik]
                    res = fc.results[ik]
                    try:
                        r = fc.device_index(stored._t.device)
                    except ValueError:
                        r = 0
                    if self._updater is not None:
                        # update on the store: the optimizer runs on the merged gradient where the weight lives; pulls
                        # then broadcast the WEIGHT (the arena holds gradient sums)
                        merged = NDArray(res[r] if stored._t.device == res[r].device else res[r].to(stored._t.device),
                                         stored.context)
                        self._updater(self._orig_key(ik), merged, stored)
                        del fc.results[ik]
                        continue
                    with torch.no_grad():
                        # device kvstore without updater: the stored value becomes the merged value (kvstore_local.h)
                        if stored._t.device == res[r].device:
                            stored._t.copy_(res[r])
                        else:
                            self._store[ik] = NDArray(res[r].clone())

    def _fast_plan(self, key, vlist):
        """Cached routing decision for the hot per-key call ``push(key, [array per GPU])`` / ``pull``: (communicator,
        permutation into its device order), or None.  Keyed by (key, device tuple); the per-call cost of the fused path is
        then a dict lookup and two list comprehensions (the python loop over ~200 keys IS the cost of a ResNet step's
        push / pull, tools/bandwidth/measure.py)."""
        devs = tuple(v._t.device.index if (type(v) is NDArray and v._t.is_cuda) else -1 for v in vlist)
        pk = (key, devs)
        plan = self._plans.get(pk, False)
        if plan is not False:
            return plan
        plan = None
        ik = self._int_key(key)
        if ik in self._store and ik <= MAX_ALLOWED_KEY_FOR_UPDATE and type(self._store[ik]) is NDArray and -1 not in devs and \
                tuple(self._store[ik].shape) == tuple(vlist[0].shape) and self._store[ik]._t.dtype == vlist[0]._t.dtype:
            fc = self._fused_comm(vlist)
            if fc is not None:
                order = sorted(range(len(vlist)), key=lambda i: fc.device_index(vlist[i]._t.device))
                plan = (fc, order, ik, tuple(vlist[0].shape), vlist[0]._t.dtype)
        self._plans[pk] = plan
        return plan

    def push(self, key, value, priority=0):
        if self._device_mode and type(value) is list and len(value) > 1 and isinstance(key, (int, str)):
            plan = self._fast_plan(key, value)
            if plan is not None:
                fc, order, ik, shape, dt = plan
                ts = [value[i]._t for i in order]
                if all(t.dtype is dt and tuple(t.shape) == shape and t.is_contiguous() for t in ts):
                    if not fc.push(ik, ts):
                        self._flush_fused()
                        fc.push(ik, ts)
                    return
        keys, vals = _key_value_lists(key, value)
        for ik, vlist in self._group(keys, vals):
            if ik not in self._store:
                raise MXNetError("key %s has not been initialized" % str(ik))
            fc = self._fused_comm(vlist) if ik <= MAX_ALLOWED_KEY_FOR_UPDATE and type(self._store[ik]) is NDArray and \
                tuple(self._store[ik].shape) == tuple(vlist[0].shape) and self._store[ik]._t.dtype == vlist[0]._t.dtype else None
            if fc is not None:
                order = sorted(vlist, key=lambda v: fc.device_index(v._t.device))
                if not fc.push(ik, [v._t for v in order]):
                    self._flush_fused()
                    fc.push(ik, [v._t for v in order])
                continue
            if self._fused:
                self._flush_fused()
            merged = self._reduce(ik, vlist)
            stored = self._store[ik]
            if ik > MAX_ALLOWED_KEY_FOR_UPDATE:
                # exclude-update (aux) key: average over contributing workers, never optimised
                merged = merged / float(self.num_workers) if self.num_workers > 1 else merged
                self._assign(stored, merged)
            elif self._updater is not None:
                if isinstance(merged, NDArray) and isinstance(stored, NDArray) and merged._t.device != stored._t.device:
                    merged = NDArray(merged._t.to(stored._t.device), stored.context)
                self._updater(self._orig_key(ik), merged, stored)
            elif (self._device_mode and isinstance(merged, NDArray) and isinstance(stored, NDArray) and merged._t.is_cuda
                  and merged._t.shape == stored._t.shape):
                # device kvstore without updater: the stored value BECOMES the merged device buffer (the reference does
                # ``local = merged`` here, kvstore_local.h PushImpl) -- no round trip through host memory
                own = all(merged is not v for v in vlist)
                self._store[ik] = merged if own else NDArray(merged._t.clone(), merged.context)
            else:
                self._assign(stored, merged)

    def _orig_key(self, ik):
        if self._key_type is str:
            for k, v in self._str2int.items():
                if v == ik:
                    return k
        return ik

    @staticmethod
    def _assign(stored, merged):
        if isinstance(stored, _sp.RowSparseNDArray):
            src = merged if isinstance(merged, _sp.RowSparseNDArray) else _sp.cast_storage(merged, "row_sparse")
            src.copyto(stored)
        elif isinstance(merged, _sp.BaseSparseNDArray):
            stored[:] = merged.todense()
        else:
            with torch.no_grad():
                stored._t.copy_(merged._t)

    def pull(self, key, out=None, priority=0, ignore_sparse=True):
        assert out is not None
        if self._fused and type(out) is list and isinstance(key, (int, str)):
            # hot per-key call: every GPU already holds the reduced value in its arena
            if any(f and f.pending for f in self._fused.values()):
                self._flush_fused()
            ik = key if isinstance(key, int) else self._str2int.get(key)
            for fc in self._fused.values():
                res = fc.results.get(ik) if fc else None
                if res is not None:
                    try:
                        pairs = [(o._t, res[fc.device_index(o._t.device)]) for o in out]
                    except (ValueError, AttributeError):
                        break
                    if all(d.dtype is s_.dtype and d.shape == s_.shape for d, s_ in pairs):
                        with torch.no_grad():
                            # one call for all GPUs: the multi-tensor copy groups its operands by device itself
                            torch._foreach_copy_([d for d, _ in pairs], [s_ for _, s_ in pairs], non_blocking=True)
                        return
                    break
        keys, outs = _key_value_lists(key, out)
        if self._fused:
            self._flush_fused()
            # keys reduced by the last flush: every GPU already holds the result in its arena -> one multi-tensor copy
            # per GPU instead of a broadcast from the owner
            per_dev = {}
            rest_k, rest_o = [], []
            for k, olist in zip(keys, outs):
                ik = self._int_key(k)
                fc = next((f for f in self._fused.values() if f and ik in f.results), None)
                ok = fc is not None and all(type(o) is NDArray and o._t.is_cuda and o._t.device in fc.devices
                                            and o._t.dtype == fc.results[ik][0].dtype for o in olist)
                if not ok:
                    rest_k.append(k); rest_o.append(olist)
                    continue
                for o in olist:
                    r = fc.device_index(o._t.device)
                    dst, src = per_dev.setdefault(o._t.device, ([], []))
                    dst.append(o._t); src.append(fc.results[ik][r])
            with torch.no_grad():
                for dev, (dst, src) in per_dev.items():
                    with torch.cuda.device(dev):
                        torch._foreach_copy_(dst, src)
            keys, outs = rest_k, rest_o
        for k, olist in zip(keys, outs):
            ik = self._int_key(k)
            if ik not in self._store:
                raise MXNetError("key %s has not been initialized" % str(k))
            stored = self._store[ik]
            if isinstance(stored, _sp.RowSparseNDArray) and ignore_sparse:
                continue                      # kvstore.py:pull -- sparse values are skipped unless ignore_sparse=False
            for o in olist:
                if isinstance(o, _sp.BaseSparseNDArray):
                    if ignore_sparse:
                        continue
                    o[:] = stored             # full row_sparse value into a row_sparse output
                    continue
                with torch.no_grad():
                    o._t.copy_(stored._t, non_blocking=True)

    def row_sparse_pull(self, key, out=None, priority=0, row_ids=None):
        assert out is not None and row_ids is not None
        keys, outs = _key_value_lists(key, out)
        _k2, rids = _key_value_lists(key, row_ids)
        from .utils import unique_rows
        for k, olist, rlist in zip(keys, outs, rids):
            ik = self._int_key(k)
            stored = self._store[ik]
            for o in olist:
                if not isinstance(o, _sp.RowSparseNDArray):
                    raise MXNetError("row_sparse_pull: the output for key %s must be a row_sparse NDArray" % str(k))
            dense = stored.todense()._t if isinstance(stored, _sp.RowSparseNDArray) else stored._t
            if len(rlist) == 1 and len(olist) > 1:
                rlist = rlist * len(olist)
            for o, rid in zip(olist, rlist):
                ids = unique_rows(rid._t.to(dense.device), num_rows=dense.shape[0])
                data = dense.index_select(0, ids)
                if isinstance(o, _sp.RowSparseNDArray):
                    dev = o._data.device
                    o._data, o._indices, o._shape = data.to(dev), ids.to(dev), tuple(dense.shape)
                else:
                    with torch.no_grad():
                        o._t.zero_()
                        o._t.index_copy_(0, ids.to(o._t.device), data.to(o._t.device))

    def set_optimizer(self, optimizer):
        self._set_updater(opt.get_updater(optimizer))


class KVStoreDevice(KVStoreLocal):
    """kvstore('device'/'nccl') when the job is one process per GPU: the store spans the process group.

    ``push``/``pull`` of individual keys fall back to an allreduce of that key across ranks followed by the
    local updater (every rank applies the identical update), which keeps the plain KVStore API usable; the
    fast path is :meth:`fused_trainer`, which gluon.Trainer and Module use to run a whole step as one fused
    kernel per dtype group."""

    def __init__(self, kv_type="device"):
        super(KVStoreDevice, self).__init__(kv_type)
        import torch.distributed as dist
        self._dist = dist
        self._trainer = None

    @property
    def rank(self):
        return self._dist.get_rank() if self._dist.is_initialized() else 0

    @property
    def num_workers(self):
        return self._dist.get_world_size() if self._dist.is_initialized() else 1

    def init(self, key, value, exclude_update=False):
        super(KVStoreDevice, self).init(key, value, exclude_update)
        if self.num_workers > 1:   # only rank 0's value is used (kvstore.py:116-158)
            keys, _ = _key_value_lists(key, value)
            for k in keys:
                st = self._store[self._int_key(k)]
                if isinstance(st, NDArray):
                    self._collective(self._dist.broadcast, st._t, src=0)

    def _reduce(self, ikey, vals):
        merged = super(KVStoreDevice, self)._reduce(ikey, vals)
        if self.num_workers > 1 and isinstance(merged, NDArray):
            t = merged._t.clone() if merged is vals[0] else merged._t
            self._collective(self._dist.all_reduce, t)
            merged = NDArray(t, merged.context)
        return merged

    def _collective(self, fn, t, **kw):
        """Run an in-place collective on ``t``; host tensors are staged through the rank's GPU when the process group
        is NCCL-only (values initialised from mx.nd.zeros(shape) live on the CPU, like the reference's pinned copy)."""
        if not t.is_cuda and torch.cuda.is_available() and self._dist.get_backend() == "nccl":
            g = t.to(torch.device("cuda", torch.cuda.current_device()))
            fn(g, **kw)
            t.copy_(g)
        else:
            fn(t, **kw)

    def _barrier(self):
        if self.num_workers > 1:
            self._dist.barrier()

    def _membership_change_barrier(self, env):
        from ..parallel import membership
        return membership.process_group_membership_barrier(self, env)


class KVStoreElasticDevice(KVStoreLocal):
    """kvstore('elastic_device'): one process per GPU on ONE box whose worker set changes at epoch boundaries.

    Membership is decided by the same scheduler protocol as the parameter-server stores (host file, removal before
    addition, dense re-ranking, newcomers launched with NEW_WORKER=1 / EPOCH_BEGIN; csrc/host/pslite.cc) but there are no
    servers in the data path: ``gluon.Trainer`` runs its step as the fused P2P allreduce+update kernel over the GPUs that
    are currently in the job.  ``_membership_change_barrier`` un-shards the optimizer state, tears the peer mapping down,
    meets the other workers in the scheduler barrier (leavers exit there), rebuilds the mapping over the new rank set and
    lets every rank adopt rank 0's weights / optimizer state / aux state (parallel/membership.py)."""

    def __init__(self, kv_type="elastic_device"):
        super(KVStoreElasticDevice, self).__init__("device")
        self._type = kv_type
        from ..parallel.membership import ElasticGroup
        self.elastic_group = ElasticGroup()
        self._trainers = []
        self._closed = False
        self.reinit_log = []

    @property
    def type(self):
        return self._type

    @property
    def rank(self):
        return self.elastic_group.rank

    @property
    def num_workers(self):
        return self.elastic_group.world

    @property
    def is_new_worker(self):
        return self.elastic_group.is_new_worker

    def peer_group(self, device):
        return self.elastic_group.peer_group(device)

    def _register_trainer(self, trainer):
        import weakref
        if not any(r() is trainer for r in self._trainers):
            self._trainers.append(weakref.ref(trainer))

    def _live_trainers(self):
        return [t for t in (r() for r in self._trainers) if t is not None]

    def _barrier(self):
        self.elastic_group.barrier()

    def _membership_change_barrier(self, env):
        """The elastic re-init (north star: drain the engine -> rebuild the peer map -> re-apply the dynamic-mini-batch
        learning-rate rule, without killing the process).  Timed: ``self.reinit_log`` gets one record per boundary."""
        import time
        from .. import engine as _engine
        t0 = time.time()
        trainers = self._live_trainers()
        eng = _engine.get()
        eng.stop()                                   # drain host-side work in flight (async checkpoints, IO prefetch)
        for t in trainers:
            t._elastic_detach()                      # average aux state, un-shard optimizer state, unmap the arenas
        t1 = time.time()
        epoch = int(env.get("EPOCH_BEGIN", 0)) if isinstance(env, dict) else 0
        old, new = self.elastic_group.membership_barrier(epoch)      # a removed worker exits in here
        t2 = time.time()
        eng.start()
        for t in trainers:
            t._on_membership_change(old, new)        # dynamic-mini-batch learning-rate rule + momentum correction (on the
            #                                          exported state: newcomers then copy the corrected values)
            t._elastic_attach()                      # new peer map, new arenas, state re-sharded, newcomers synced
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        t3 = time.time()
        rec = {"epoch": epoch, "old_world": old, "new_world": new, "drain_detach_ms": (t1 - t0) * 1e3,
               "scheduler_barrier_ms": (t2 - t1) * 1e3, "attach_ms": (t3 - t2) * 1e3,
               "reinit_ms_excl_scheduler": ((t1 - t0) + (t3 - t2)) * 1e3, "total_ms": (t3 - t0) * 1e3}
        self.reinit_log.append(rec)
        return new

    def close(self):
        if not self._closed:
            self._closed = True
            for t in self._live_trainers():
                if t._fused is not None:
                    t._unfuse()
            self.elastic_group.close()


def _maybe_join_torchrun_job(dist):
    """One process per GPU is this framework's multi-GPU model (the reference drives all GPUs of a box from one process).
    A script started by ``torchrun`` / ``python -m torch.distributed.run`` (RANK, WORLD_SIZE, LOCAL_RANK, MASTER_* in
    the environment) joins the NCCL process group on its own GPU the first time it asks for a device store, so reference
    scripts need no torch-specific set-up code."""
    import os
    try:
        world = int(os.environ.get("WORLD_SIZE", "1"))
    except ValueError:
        return
    if world <= 1 or "RANK" not in os.environ or not dist.is_available() or dist.is_initialized():
        return
    if not getenv("MXB200_AUTO_PROCESS_GROUP", True):
        return
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if torch.cuda.is_available():
        local = int(os.environ.get("LOCAL_RANK", os.environ["RANK"])) % max(1, torch.cuda.device_count())
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", rank=int(os.environ["RANK"]), world_size=world, device_id=torch.device("cuda", local))
    else:
        dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=world)


def create(name="local"):
    """mx.kv.create (python/mxnet/kvstore.py:645-688; type parsing as in src/kvstore/kvstore.cc:40-77)."""
    if not isinstance(name, str):
        raise TypeError("name must be a string")
    if name.startswith("elastic_device") or ("dist" in name and "device" in name and getenv("MXB200_SINGLE_BOX_ELASTIC", False)):
        # single-box elastic job: GPUs exchange gradients over NVLink, the scheduler only manages membership
        return KVStoreElasticDevice("elastic_device")
    if "dist" in name:
        from .dist import KVStoreDist
        return KVStoreDist(name)
    if "device" in name or "nccl" in name:
        import torch.distributed as dist
        _maybe_join_torchrun_job(dist)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            return KVStoreDevice(name)
        return KVStoreLocal(name)
    if name.startswith("local"):
        return KVStoreLocal(name)
    raise MXNetError("Unknown KVStore type \"%s\"" % name)
