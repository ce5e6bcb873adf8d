"""Distributed (parameter-server) KVStore: worker side + server logic
(reference: src/kvstore/kvstore_dist.h:44-780 worker, kvstore_dist_server.h:154-797 server,
python/mxnet/kvstore_server.py).

Transport / membership are native (csrc/host/pslite.cc).  Semantics kept:
  * keys < bigarray_bound live on server (key*9973) % S; bigger values are split evenly over all servers
    (EncodeDefaultKey, kvstore_dist.h:547-589);
  * sync mode: a server applies a key when pushes from ``num_workers`` (the LIVE value) have arrived, runs the
    optimizer for ordinary keys and AVERAGES keys above MAX_ALLOWED_KEY_FOR_UPDATE (aux states)
    (kvstore_dist_server.h:345-379), then answers every queued push;
  * ``init`` uses rank 0's value; workers started with NEW_WORKER=1 skip the start-up and post-init barriers
    (kvstore_dist.h:46-61,205-224) and obtain state by pulling;
  * ``_membership_change_barrier(env)`` is the epoch-boundary barrier at which the scheduler applies host-file
    changes; afterwards ``rank`` / ``num_workers`` reflect the new membership.
"""
import ctypes
import os
import pickle
import threading

import numpy as np
import torch

from .. import _native as N
from ..base import MXNetError, getenv
from ..ndarray import NDArray
from ..ndarray.serialization import FLAG2TORCH, TORCH2FLAG
from .kvstore import KVStoreLocal, _key_value_lists, MAX_ALLOWED_KEY_FOR_UPDATE
from .. import optimizer as opt

c_void_p, c_int, c_int64, c_uint64, c_char_p = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_uint64, ctypes.c_char_p

ROLE = {"scheduler": 0, "server": 1, "worker": 2}
GROUP_ALL, GROUP_WORKER, GROUP_SERVER = 7, 4, 2
# server commands (kvstore_dist_server.h:45-48)
K_CONTROLLER, K_SET_MULTI_PRECISION, K_STOP_SERVER, K_SYNC_MODE, K_SET_GRAD_COMPRESSION, K_SET_PROFILER = 0, 1, 2, 3, 4, 5
REQ_DEFAULT, REQ_ROW_SPARSE, REQ_COMPRESSED = 0, 1, 2

_DATA_CB = ctypes.CFUNCTYPE(None, c_void_p, c_int, c_int, c_int64, c_int, c_int64, c_int64, c_void_p, c_uint64, c_int, c_int)
_CMD_CB = ctypes.CFUNCTYPE(None, c_void_p, c_int, c_char_p, c_int, c_int)
_LAUNCH_CB = ctypes.CFUNCTYPE(c_int, c_void_p, c_char_p, c_char_p)


class PsNode(object):
    """ctypes wrapper around the native node."""

    def __init__(self, role, root_uri=None, root_port=None, num_workers=None, num_servers=None, tag=None,
                 is_new_worker=None, host_file=None):
        self.lib = N.host()
        self.role = role
        env = os.environ
        root_uri = root_uri or env.get("DMLC_PS_ROOT_URI", "127.0.0.1")
        root_port = int(root_port or env.get("DMLC_PS_ROOT_PORT", "9091"))
        num_workers = int(num_workers if num_workers is not None else env.get("DMLC_NUM_WORKER", "1"))
        num_servers = int(num_servers if num_servers is not None else env.get("DMLC_NUM_SERVER", "1"))
        tag = tag if tag is not None else env.get("MXB200_WORKER_TAG", env.get("DMLC_NODE_HOST", ""))
        if is_new_worker is None:
            is_new_worker = env.get("NEW_WORKER", "0") == "1"
        host_file = host_file if host_file is not None else env.get("WORKER_HOST_FILE", "")
        self.h = c_void_p()
        N.check(self.lib, self.lib.MXBPsCreate(c_int(ROLE[role]), root_uri.encode(), c_int(root_port), c_int(num_workers),
                                               c_int(num_servers), tag.encode(), c_int(1 if is_new_worker else 0),
                                               host_file.encode(), c_int(num_workers), ctypes.byref(self.h)))
        self.is_new_worker = bool(is_new_worker)
        self._cbs = []

    def start(self):
        N.check(self.lib, self.lib.MXBPsStart(self.h))

    def stop(self):
        if self.h:
            self.lib.MXBPsStop(self.h)

    def info(self):
        a, b, c, d, e = c_int(), c_int(), c_int(), c_int(), c_int()
        N.check(self.lib, self.lib.MXBPsInfo(self.h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c), ctypes.byref(d), ctypes.byref(e)))
        return {"id": a.value, "rank": b.value, "num_workers": c.value, "num_servers": d.value, "terminated": bool(e.value)}

    def server_ids(self):
        arr = (c_int * 64)()
        n = c_int()
        N.check(self.lib, self.lib.MXBPsServerIds(self.h, arr, c_int(64), ctypes.byref(n)))
        return [arr[i] for i in range(n.value)]

    def barrier(self, group=GROUP_ALL):
        N.check(self.lib, self.lib.MXBPsBarrier(self.h, c_int(group)))

    def membership_change_barrier(self, env):
        kv = "\n".join("%s\n%s" % (k, v) for k, v in env.items())
        N.check(self.lib, self.lib.MXBPsMembershipChangeBarrier(self.h, kv.encode()))

    def allgather(self, blob):
        """All-gather one small bytes blob per worker through the scheduler; returns the list in rank order."""
        cap = 1 << 20
        out = ctypes.create_string_buffer(cap)
        n = c_int()
        N.check(self.lib, self.lib.MXBPsAllGather(self.h, bytes(blob).hex().encode(), out, c_int(cap), ctypes.byref(n)))
        if n.value >= cap:
            raise MXNetError("allgather result too large")
        parts = out.value.decode().split("\n")
        return [bytes.fromhex(p) for p in parts if p != ""]

    def new_request(self, expected):
        ts = c_int()
        N.check(self.lib, self.lib.MXBPsNewRequest(self.h, c_int(expected), ctypes.byref(ts)))
        return ts.value

    def send_data(self, recver, ts, push, key, dtype_flag, offset, total, head, tensor, pull_dst=None):
        ptr = c_void_p(tensor.data_ptr()) if tensor is not None else c_void_p(0)
        nbytes = tensor.numel() * tensor.element_size() if tensor is not None else 0
        dst = c_void_p(pull_dst.data_ptr()) if pull_dst is not None else c_void_p(0)
        esz = pull_dst.element_size() if pull_dst is not None else 0
        N.check(self.lib, self.lib.MXBPsSendData(self.h, c_int(recver), c_int(ts), c_int(1 if push else 0), c_int64(key),
                                                 c_int(dtype_flag), c_int64(offset), c_int64(total), c_int(head), ptr,
                                                 c_uint64(nbytes), dst, c_uint64(esz)))

    def wait(self, ts):
        N.check(self.lib, self.lib.MXBPsWait(self.h, c_int(ts)))

    def respond(self, recver, ts, key, tensor=None):
        ptr = c_void_p(tensor.data_ptr()) if tensor is not None else c_void_p(0)
        nbytes = tensor.numel() * tensor.element_size() if tensor is not None else 0
        N.check(self.lib, self.lib.MXBPsRespond(self.h, c_int(recver), c_int(ts), c_int64(key), ptr, c_uint64(nbytes)))

    def send_command(self, head, body):
        if isinstance(body, str):
            body = body.encode("latin-1")
        N.check(self.lib, self.lib.MXBPsSendCommand(self.h, c_int(head), body))

    def respond_command(self, recver, ts):
        N.check(self.lib, self.lib.MXBPsRespondCommand(self.h, c_int(recver), c_int(ts)))

    def set_data_handler(self, fn):
        cb = _DATA_CB(fn)
        self._cbs.append(cb)
        N.check(self.lib, self.lib.MXBPsSetDataHandler(self.h, cb, None))

    def set_command_handler(self, fn):
        cb = _CMD_CB(fn)
        self._cbs.append(cb)
        N.check(self.lib, self.lib.MXBPsSetCommandHandler(self.h, cb, None))

    def set_launch_callback(self, fn):
        cb = _LAUNCH_CB(fn)
        self._cbs.append(cb)
        N.check(self.lib, self.lib.MXBPsSetLaunchCallback(self.h, cb, None))

    def num_dead_nodes(self, timeout=60):
        n = c_int()
        N.check(self.lib, self.lib.MXBPsNumDeadNodes(self.h, c_int(timeout), ctypes.byref(n)))
        return n.value


def _slices(ikey, total, servers, bigarray_bound):
    """[(server_id, begin, end)] for a value of ``total`` elements (kvstore_dist.h:547-589)."""
    S = len(servers)
    if total < bigarray_bound or S == 1:
        return [(servers[(ikey * 9973) % S], 0, total)]
    out = []
    for i in range(S):
        b, e = int(round(total / S * i)), int(round(total / S * (i + 1)))
        if e > b:
            out.append((servers[i], b, e))
    return out


class KVStoreDist(KVStoreLocal):
    """Worker-side distributed store (dist_sync / dist_device_sync / dist_async)."""

    def __init__(self, kv_type="dist_sync"):
        super(KVStoreDist, self).__init__(kv_type)
        role = os.environ.get("DMLC_ROLE", "worker")
        if role != "worker":
            raise MXNetError("KVStoreDist can only be created in a worker process (DMLC_ROLE=%s)" % role)
        self._node = PsNode("worker")
        self._node.start()
        self._bigarray_bound = int(getenv("MXNET_KVSTORE_BIGARRAY_BOUND", 1000000))
        self._is_new_worker = self._node.is_new_worker
        self._keys_shape = {}
        self._closed = False
        if not self._is_new_worker:
            self._node.barrier(GROUP_ALL)
            if self.rank == 0:
                self._node.send_command(K_SYNC_MODE, "1" if "_async" not in kv_type else "0")
            self._node.barrier(GROUP_WORKER)
        import atexit
        atexit.register(self.close)

    # identity ------------------------------------------------------------------------------------
    @property
    def rank(self):
        return self._node.info()["rank"]

    @property
    def num_workers(self):
        return self._node.info()["num_workers"]

    def get_num_dead_node(self, node_id=0, timeout=60):
        return self._node.num_dead_nodes(timeout)

    # data ----------------------------------------------------------------------------------------
    def _host(self, nd):
        t = nd._t.detach()
        return t.to("cpu").contiguous() if t.is_cuda or not t.is_contiguous() else t

    def _push_tensor(self, ik, t, head=REQ_DEFAULT, total=None):
        servers = self._node.server_ids()
        total = t.numel() if total is None else total
        if head == REQ_COMPRESSED:
            sl = [(servers[(ik * 9973) % len(servers)], 0, t.numel())]
        else:
            sl = _slices(ik, t.numel(), servers, self._bigarray_bound)
        ts = self._node.new_request(len(sl))
        flat = t.reshape(-1)
        keep = []
        for sid, b, e in sl:
            part = flat[b:e].contiguous()
            keep.append(part)
            self._node.send_data(sid, ts, True, ik, TORCH2FLAG[t.dtype], b, total, head, part)
        self._node.wait(ts)

    def init(self, key, value, exclude_update=False):
        keys, vals = _key_value_lists(key, value)
        for k, vlist in zip(keys, vals):
            ik = self._int_key(k, create=True, exclude_update=exclude_update)
            if ik in self._keys_shape:
                raise MXNetError("duplicate init of key %s" % str(k))
            v = vlist[0]
            self._keys_shape[ik] = (tuple(v.shape), v._t.dtype)
            self._store[ik] = v.copy()
            if self.rank == 0 and not self._is_new_worker:
                self._push_tensor(ik, self._host(v))       # first push of a key == init on the server
        if not self._is_new_worker:
            self._node.barrier(GROUP_WORKER)

    def push(self, key, value, priority=0):
        keys, vals = _key_value_lists(key, value)
        for ik, vlist in self._group(keys, vals):
            if ik not in self._keys_shape:
                raise MXNetError("key %s has not been initialized" % str(ik))
            merged = self._reduce(ik, vlist) if len(vlist) > 1 else vlist[0]
            t = self._host(merged)
            if self._compression is not None and ik <= MAX_ALLOWED_KEY_FOR_UPDATE:
                res = self._residuals.setdefault((ik, "dist"), torch.zeros(t.shape, dtype=torch.float32))
                words = self._compression.quantize(t, res)
                self._push_tensor(ik, words, head=REQ_COMPRESSED, total=t.numel())
            else:
                self._push_tensor(ik, t)

    def pull(self, key, out=None, priority=0, ignore_sparse=True):
        assert out is not None
        keys, outs = _key_value_lists(key, out)
        for k, olist in zip(keys, outs):
            ik = self._int_key(k)
            if ik not in self._keys_shape:
                raise MXNetError("key %s has not been initialized" % str(k))
            shape, dtype = self._keys_shape[ik]
            recv = torch.empty(int(np.prod(shape)), dtype=dtype)
            servers = self._node.server_ids()
            sl = _slices(ik, recv.numel(), servers, self._bigarray_bound)
            ts = self._node.new_request(len(sl))
            for sid, b, e in sl:
                self._node.send_data(sid, ts, False, ik, TORCH2FLAG[dtype], b, recv.numel(), REQ_DEFAULT, None,
                                     pull_dst=recv[b:e])
            self._node.wait(ts)
            recv = recv.reshape(shape)
            for o in olist:
                with torch.no_grad():
                    o._t.copy_(recv.to(o._t.device), non_blocking=True)

    def row_sparse_pull(self, key, out=None, priority=0, row_ids=None):
        keys, outs = _key_value_lists(key, out)
        _k, rids = _key_value_lists(key, row_ids)
        from .utils import unique_rows
        from ..ndarray import sparse as _sp
        for k, olist, rlist in zip(keys, outs, rids):
            ik = self._int_key(k)
            shape, dtype = self._keys_shape[ik]
            dense = NDArray(torch.empty(shape, dtype=dtype))
            self.pull(k, out=dense)
            if len(rlist) == 1 and len(olist) > 1:
                rlist = rlist * len(olist)
            for o, rid in zip(olist, rlist):
                ids = unique_rows(rid._t.cpu())
                data = dense._t.index_select(0, ids)
                if isinstance(o, _sp.RowSparseNDArray):
                    dev = o._data.device
                    o._data, o._indices, o._shape = data.to(dev), ids.to(dev), tuple(shape)
                else:
                    with torch.no_grad():
                        o._t.zero_()
                        o._t.index_copy_(0, ids.to(o._t.device), data.to(o._t.device))

    # control -------------------------------------------------------------------------------------
    def set_optimizer(self, optimizer):
        """Pickle the optimizer to the servers (kController) -- worker rank 0 only (kvstore.py:451-498)."""
        if self.rank == 0 and not self._is_new_worker:
            body = pickle.dumps(optimizer, 0).decode("latin-1")
            self._node.send_command(K_CONTROLLER, body)
            if optimizer.multi_precision:
                self._node.send_command(K_SET_MULTI_PRECISION, "")
        if not self._is_new_worker:
            self._node.barrier(GROUP_WORKER)

    def set_gradient_compression(self, compression_params):
        super(KVStoreDist, self).set_gradient_compression(compression_params)
        if self.rank == 0 and not self._is_new_worker:
            self._node.send_command(K_SET_GRAD_COMPRESSION, self._compression.encode_params())
        if not self._is_new_worker:
            self._node.barrier(GROUP_WORKER)

    def _send_command_to_servers(self, head, body):
        self._node.send_command(head, body)

    def _barrier(self):
        self._node.barrier(GROUP_WORKER)

    def _membership_change_barrier(self, env):
        self._node.membership_change_barrier({str(k): str(v) for k, v in env.items()})
        self._is_new_worker = False
        if self._node.info()["terminated"]:
            raise SystemExit(0)       # this worker was removed from the job at the epoch boundary

    def save_optimizer_states(self, fname, dump_optimizer=False):
        raise MXNetError("Cannot save states for distributed training")

    def load_optimizer_states(self, fname):
        raise MXNetError("Cannot load states for distributed training")

    def close(self):
        if self._closed:
            return
        self._closed = True
        try:
            info = self._node.info()
            if not info["terminated"]:
                self._node.barrier(GROUP_WORKER)
                if info["rank"] == 0:
                    self._node.send_command(K_STOP_SERVER, "")
                self._node.barrier(GROUP_ALL)
        finally:
            self._node.stop()


class KVStoreDistServer(object):
    """Server-side storage + aggregation + optimizer (kvstore_dist_server.h)."""

    def __init__(self, node):
        self.node = node
        self.store = {}          # key -> {offset: tensor slice}
        self.merge = {}          # (key, offset) -> (accumulated tensor, [(sender, ts)])
        self.sync_mode = True
        self.updater = None
        self.compression = None
        self.stop_event = threading.Event()
        self.lock = threading.Lock()
        node.set_data_handler(self._on_data)
        node.set_command_handler(self._on_command)

    def _on_command(self, _user, head, body, sender, ts):
        body = body.decode("latin-1") if body else ""
        if head == K_CONTROLLER:
            optimizer = pickle.loads(body.encode("latin-1"))
            self.updater = opt.get_updater(optimizer)
        elif head == K_SYNC_MODE:
            self.sync_mode = body != "0"
        elif head == K_SET_GRAD_COMPRESSION:
            from .compression import GradientCompression
            typ, thr = body.split(",")
            self.compression = GradientCompression({"type": "2bit" if typ == "1" else "none", "threshold": float(thr)})
        elif head == K_SET_PROFILER:
            from .. import profiler
            import json
            try:
                req = json.loads(body)
                info = self.node.info()
                if req["cmd"] == "set_config":
                    args = dict(req["args"])
                    args.pop("profile_process", None)
                    if "filename" in args:
                        args["filename"] = "rank%d_%s" % (info["rank"], args["filename"])
                    profiler.set_config(**args)
                elif req["cmd"] == "set_state":
                    profiler.set_state(req["args"])
            except Exception:  # noqa: BLE001
                pass
        self.node.respond_command(sender, ts)
        if head == K_STOP_SERVER:
            self.stop_event.set()

    def _apply(self, key, offset, merged):
        stored = self.store[key][offset]
        nw = self.node.info()["num_workers"]
        if key > MAX_ALLOWED_KEY_FOR_UPDATE:
            stored.copy_((merged / float(nw)).to(stored.dtype))      # aux state: average over live workers
        elif self.updater is not None:
            w, g = NDArray(stored), NDArray(merged.to(stored.dtype))
            self.updater(key, g, w)     # each server owns at most one slice of a key, so the key indexes the state
        else:
            stored.copy_(merged.to(stored.dtype))

    def _on_data(self, _user, is_push, sender, key, dtype, offset, total, data, nbytes, ts, head):
        with self.lock:
            if is_push:
                if head == REQ_COMPRESSED:
                    words = torch.frombuffer(bytearray(ctypes.string_at(data, nbytes)), dtype=torch.int32)
                    t = self.compression.dequantize(words, int(total), (int(total),))
                else:
                    td = FLAG2TORCH[dtype]
                    t = torch.frombuffer(bytearray(ctypes.string_at(data, nbytes)), dtype=td).clone()
                slot = self.store.setdefault(key, {})
                if offset not in slot:
                    slot[offset] = t.clone()                           # first push == init
                    self.node.respond(sender, ts, key)
                    return
                if not self.sync_mode:
                    self._apply(key, offset, t.float())
                    self.node.respond(sender, ts, key)
                    return
                acc, waiting = self.merge.get((key, offset), (None, []))
                acc = t.float() if acc is None else acc + t.float()
                waiting.append((sender, ts))
                if len(waiting) >= self.node.info()["num_workers"]:
                    self._apply(key, offset, acc)
                    self.merge.pop((key, offset), None)
                    for s, tts in waiting:
                        self.node.respond(s, tts, key)
                else:
                    self.merge[(key, offset)] = (acc, waiting)
            else:
                slot = self.store.get(key, {})
                if offset not in slot:
                    self.node.respond(sender, ts, key, torch.zeros(0))
                    return
                self.node.respond(sender, ts, key, slot[offset])

    def run(self):
        self.node.barrier(GROUP_ALL)
        self.stop_event.wait()
        self.node.barrier(GROUP_ALL)
        self.node.stop()
