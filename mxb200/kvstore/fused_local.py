"""Bucketed per-key push / pull on the fused kernel for the reference's process model: ONE process, one array per GPU.

Reference path being replaced: for every key, CommDevice::Reduce = (N-1) peer copies + ElementwiseSum on a greedily chosen
owner GPU, then CommDevice::Broadcast = N copies back (src/kvstore/comm.h:503-616; python loop in model.py:150-170 /
gluon/trainer.py:_allreduce_grads) -- ~205 x 2 x N small operations per ResNet-200 step, 4.4 GB/s per GPU in the
reference's own measurement (tools/bandwidth/README.md:50-57).

Here ``push`` only records the request (the reference's push is asynchronous too: an engine op with a priority).  The
first ``pull`` (or ``waitall`` / a push of a key that is already pending) flushes ALL pending keys at once:

  * per GPU one multi-tensor copy packs that GPU's values into its slot of a symmetric, peer-mapped arena
    (layout: parallel/arena.py, one arena per dtype, built over the keys known at the first flush);
  * per GPU ONE launch of the fused kernel in plain-sum mode (csrc/cuda/allreduce.cu kAllreduce): reduce-scatter over
    peer memory + all-gather by peer stores (or in-switch multimem), flags instead of host synchronisation;
  * results stay in each GPU's arena; ``pull`` copies the key's slice into the caller's arrays with one multi-tensor
    copy per GPU.  The store's own value of the key (the merged array) is updated on its device.

Falls back (returns False) for sparse values, gradient compression, an updater on the store, CPU arrays, GPUs without
peer access or a single device.
"""
import torch

from ..base import getenv
from ..ndarray import NDArray

__all__ = ["FusedLocalComm"]


class FusedLocalComm(object):
    def __init__(self, devices):
        from ..parallel import peer
        self.devices = [torch.device(d) for d in devices]
        self.groups = peer.make_local_groups(self.devices)
        self.arenas = {}            # dtype -> [SymmetricArena per device]
        self.layout_keys = {}       # dtype -> {ikey: index in layout}
        self.pending = {}           # ikey -> [tensor per device] (device order)
        self.results = {}           # ikey -> [arena view per device]   valid until the next flush
        self.launches = 0
        self.seq = 0

    @staticmethod
    def usable(vals):
        if not getenv("MXB200_KV_FUSED", True) or len(vals) < 2:
            return None
        devs = []
        for v in vals:
            if not isinstance(v, NDArray) or type(v) is not NDArray:
                return None
            t = v._t
            if not t.is_cuda or t.dtype not in (torch.float32, torch.bfloat16) or not t.is_contiguous():
                return None
            devs.append(t.device)
        if len(set(devs)) != len(devs):
            return None
        return devs

    def close(self):
        for ars in self.arenas.values():
            ars[:] = []
        for g in self.groups:
            g.close()

    # ------------------------------------------------------------------------------------------------
    def _build(self, dtype, store):
        """(Re)build the arena of one dtype over every dense key of that dtype in the store."""
        from ..parallel import fused
        from ..parallel.peer import _LocalGroup
        keys = sorted(ik for ik, v in store.items() if type(v) is NDArray and v._t.dtype is dtype)
        tmpl = [store[ik]._t for ik in keys]
        names = [str(ik) for ik in keys]
        nbytes = fused.SymmetricArena.bytes_needed(tmpl, dtype)
        wb = _LocalGroup.alloc_all(self.groups, nbytes)
        gb = _LocalGroup.alloc_all(self.groups, nbytes)
        ars = []
        for r, g in enumerate(self.groups):
            with torch.cuda.device(g.device):
                shapes = [torch.empty(t.shape, dtype=dtype, device="meta") for t in tmpl]
                ar = fused.SymmetricArena.__new__(fused.SymmetricArena)
                self._init_arena(ar, g, names, shapes, {"weights": wb[r], "grads": gb[r]}, dtype)
            ars.append(ar)
        self.arenas[dtype] = ars
        self.layout_keys[dtype] = {ik: i for i, ik in enumerate(keys)}

    @staticmethod
    def _init_arena(ar, group, names, shapes, bufs, dtype):
        """SymmetricArena without copying template values in (the kv arena is a pure communication buffer)."""
        from ..parallel.arena import ArenaLayout
        from ..parallel.peer import MAX_RANKS
        ar.group, ar.device = group, group.device
        ar.layout = ArenaLayout(names, [t.shape for t in shapes], align_chunks_to=MAX_RANKS)
        ar.n = ar.layout.total
        ar.wdtype = ar.gdtype = dtype
        (wl, ar.w_ptrs), (gl, ar.g_ptrs) = bufs["weights"], bufs["grads"]
        ar.w_mc = ar.g_mc = None
        ar._one_shot = None
        ar.bw_epoch, ar.dirty, ar.write_semantics = -1, False, False
        ar.norm_bufs = ar.norm_tot = ar.gsync = None
        ar.ll_bufs = None
        ar.weights = wl.tensor(dtype, ar.n)
        ar.grads = gl.tensor(dtype, ar.n)
        ar.chunk_tid = torch.from_numpy(ar.layout.chunk_tid).to(ar.device)
        ar.lr_mult = torch.ones(len(names), dtype=torch.float32, device=ar.device)
        ar.wd_mult = torch.ones(len(names), dtype=torch.float32, device=ar.device)
        ar.state = {}
        ar.views = [ar.layout.view(ar.weights, i, False) for i in range(len(names))]
        ar.grad_views = [ar.layout.view(ar.grads, i, False) for i in range(len(names))]

    # ------------------------------------------------------------------------------------------------
    def push(self, ikey, tensors_by_device):
        """Record a push (tensors in self.devices order).  A second push of a pending key flushes first."""
        if ikey in self.pending:
            return False
        self.pending[ikey] = tensors_by_device
        return True

    def flush(self, store):
        """Reduce every pending key; afterwards ``self.results[ikey][r]`` is the sum as seen by device r."""
        from ..parallel import fused
        if not self.pending:
            return
        pend, self.pending = self.pending, {}
        self.results = {}
        by_dtype = {}
        for ik, ts in pend.items():
            by_dtype.setdefault(ts[0].dtype, []).append(ik)
        for dtype, iks in by_dtype.items():
            lk = self.layout_keys.get(dtype)
            if lk is None or any(ik not in lk for ik in iks) or \
                    any(tuple(pend[ik][0].shape) != self.arenas[dtype][0].layout.shapes[lk[ik]] for ik in iks):
                self._build(dtype, store)
                lk = self.layout_keys[dtype]
            ars = self.arenas[dtype]
            with torch.no_grad():
                for r, ar in enumerate(ars):
                    with torch.cuda.device(ar.device):
                        torch._foreach_copy_([ar.grad_views[lk[ik]] for ik in iks], [pend[ik][r] for ik in iks])
                self.seq += 1
                for r, ar in enumerate(ars):
                    with torch.cuda.device(ar.device):
                        # every GPU: ONE launch; the CTAs of the N kernels meet in the flag barriers over peer memory
                        fused.fused_step(ar, fused.MODE_ALLREDUCE, rescale=1.0, zero_grad=False, seq=self.seq)
                        self.launches += 1
            for ik in iks:
                self.results[ik] = [ar.views[lk[ik]] for ar in ars]

    def device_index(self, device):
        return self.devices.index(torch.device(device))
