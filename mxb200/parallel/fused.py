"""Launcher for the fused allreduce + optimizer kernel (csrc/cuda/allreduce.cu).

``SymmetricArena`` owns one dtype-group of parameters on one rank: symmetric (peer-mapped) weight and
gradient buffers, local fp32 master / optimizer state, chunk table.  ``fused_step`` is the whole
``kv.push(grads) -> optimizer -> kv.pull(weights)`` sequence of the reference
(python/mxnet/model.py:183-205 -> kvstore_local.h:189-242) as one kernel launch per rank.
"""
import ctypes

import numpy as np
import torch

from .. import _native as N
from ..ops._dtypes import code
from .arena import ArenaLayout, CHUNK
from .peer import MAX_RANKS

c_void_p, c_int, c_int64, c_float, c_uint32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_uint32

LAUNCHES = [0]
_UNROLL = [None]

MODE_ALLREDUCE, MODE_SGD, MODE_ADAM, MODE_LAMB1, MODE_LAMB2, MODE_BCAST, MODE_LAMB = 0, 1, 2, 3, 4, 5, 6

# one-shot below this many bytes of gradient per rank (latency-bound), two-shot above
ONE_SHOT_BYTES = 512 * 1024
# ... and below this the barrier-free low-latency protocol (flag-carrying 8-byte packets pushed into every peer's buffer)
LL_BYTES = 256 * 1024


class _Args(ctypes.Structure):
    _fields_ = [("grads", c_void_p * MAX_RANKS), ("weights", c_void_p * MAX_RANKS), ("flags", c_void_p * MAX_RANKS),
                ("w32", c_void_p), ("mom", c_void_p), ("var", c_void_p), ("rbuf", c_void_p), ("w_norm2", c_void_p),
                ("r_norm2", c_void_p), ("chunk_tid", c_void_p), ("lr_mult", c_void_p), ("wd_mult", c_void_p),
                ("err", c_void_p), ("n", c_int64), ("rank", c_int), ("nranks", c_int), ("mode", c_int),
                ("one_shot", c_int), ("g_dt", c_int), ("w_dt", c_int), ("seq", c_uint32), ("zero_grad", c_int),
                ("lr", c_float), ("wd", c_float), ("momentum", c_float), ("rescale", c_float), ("clip", c_float),
                ("nesterov", c_int), ("has_mom", c_int), ("beta1", c_float), ("beta2", c_float), ("eps", c_float),
                ("bc1", c_float), ("bc2", c_float), ("lamb_lower", c_float), ("lamb_upper", c_float), ("hyper", c_void_p),
                ("grads_mc", c_void_p), ("weights_mc", c_void_p), ("norms", c_void_p * MAX_RANKS), ("norm_tot", c_void_p),
                ("gsync", c_void_p), ("nt_pad", c_int), ("pad0", c_int), ("state_off", c_int64), ("seq_dev", c_void_p),
                ("ll_recv", c_void_p * MAX_RANKS), ("ll_ctr", c_void_p)]

HYPER_FLOATS = 8      # csrc/cuda/allreduce.cu: [lr, wd, momentum, rescale, clip, seq(uint32 bits), bc1, bc2]


def pack_hyper(lr=0.0, wd=0.0, momentum=0.0, rescale=1.0, clip=None, seq=0, bc1=1.0, bc2=1.0):
    """Host image of one hyper-parameter record (see ``Args.hyper``): the per-step scalars of a fused launch that a
    CUDA graph must NOT freeze."""
    a = np.zeros(HYPER_FLOATS, dtype=np.float32)
    a[0], a[1], a[2], a[3] = lr, wd, momentum, rescale
    a[4] = -1.0 if clip is None else clip
    a[5:6].view(np.uint32)[0] = np.uint32(seq & 0xFFFFFFFF)
    a[6], a[7] = bc1, bc2
    return a


def _check_abi(lib):
    n = lib.MXBFusedAllreduceArgsSize()
    if n != ctypes.sizeof(_Args):
        raise RuntimeError("fused allreduce ABI mismatch: native %d bytes vs python %d" % (n, ctypes.sizeof(_Args)))


class SymmetricArena(object):
    """One dtype group on one rank.  ``group`` is a PeerGroup/_LocalGroup; buffers come from ``bufs``:
    dict with 'weights' and 'grads' -> (local _RawBuffer, [ptr per rank])."""

    def __init__(self, group, names, tensors, bufs, wdtype, gdtype):
        self.group = group
        self.device = group.device
        self.layout = ArenaLayout(names, [t.shape for t in tensors], align_chunks_to=MAX_RANKS)
        n = self.layout.total
        self.n = n
        self.wdtype, self.gdtype = wdtype, gdtype
        (wl, self.w_ptrs), (gl, self.g_ptrs) = bufs["weights"], bufs["grads"]
        # NVLS: multicast mappings of the two arenas (None -> P2P loads / stores over the unicast peer mappings)
        self.w_mc, self.g_mc = bufs.get("weights_mc"), bufs.get("grads_mc")
        self._one_shot = None
        self.ll_bufs = None               # low-latency protocol receive buffers (local buffer, [ptr per rank])
        self.bw_epoch, self.dirty, self.write_semantics = -1, False, True     # ops.functional.sink_touch
        self.norm_bufs = None             # fused LAMB: (local buffer, [ptr per rank]) of per-tensor partial norms
        self.norm_tot = self.gsync = None
        self.weights = wl.tensor(wdtype, n)
        self.grads = gl.tensor(gdtype, n)
        with torch.no_grad():
            # raw cudaMalloc memory (possibly recycled from a closed trainer): the kernel reads whole chunks, including the
            # alignment gaps, and a trainer built before the first backward pass must start from zero gradients
            self.weights.zero_()
            self.grads.zero_()
        self.chunk_tid = torch.from_numpy(self.layout.chunk_tid).to(self.device)
        self.lr_mult = torch.ones(len(names), dtype=torch.float32, device=self.device)
        self.wd_mult = torch.ones(len(names), dtype=torch.float32, device=self.device)
        self.state = {}
        self.views, self.grad_views = [], []
        with torch.no_grad():
            for i, t in enumerate(tensors):
                krsc = t.dim() == 4 and not t.is_contiguous() and t.permute(0, 2, 3, 1).is_contiguous()
                v = self.layout.view(self.weights, i, krsc)
                v.copy_(t)
                self.views.append(v)
                self.grad_views.append(self.layout.view(self.grads, i, krsc))

    @staticmethod
    def bytes_needed(tensors, dtype):
        lay = ArenaLayout([str(i) for i in range(len(tensors))], [t.shape for t in tensors], align_chunks_to=MAX_RANKS)
        return lay.total * torch.empty((), dtype=dtype).element_size()

    def one_shot(self):
        """One-shot (every rank reduces the whole message) below ONE_SHOT_BYTES of gradient, two-shot above; a caller
        may pin the choice with ``fused_step(one_shot=...)`` before any optimizer state exists (it fixes the sharding)."""
        if self._one_shot is not None:
            return self._one_shot
        return self.n * self.grads.element_size() <= ONE_SHOT_BYTES

    def slice_range(self, rank=None, world=None):
        """Arena element range [lo, hi) whose optimizer state lives on ``rank``: the slice the two-shot kernel assigns to it
        (csrc/cuda/allreduce.cu: per = ceil(n_chunks / world) chunks of 1024 elements).  One-shot arenas and single-rank
        groups keep (and update) the full state on every rank."""
        rank = self.group.rank if rank is None else rank
        world = self.group.world if world is None else world
        if world == 1 or self.one_shot():
            return 0, self.n
        n_chunks = self.n // CHUNK
        per = (n_chunks + world - 1) // world
        lo = min(n_chunks, per * rank)
        return lo * CHUNK, min(n_chunks, lo + per) * CHUNK

    def get_state(self, name, init_from_weights=False):
        """fp32 optimizer state (``w32`` master weights, ``mom``, ``var``, ``rbuf``) for the OWNED slice only: with N ranks
        every rank holds 1/N of the master weights and moments (ZeRO-1 memory, not just ZeRO-1 compute)."""
        st = self.state.get(name)
        if st is None:
            lo, hi = self.slice_range()
            st = torch.zeros(max(hi - lo, 8), dtype=torch.float32, device=self.device)
            if init_from_weights and hi > lo:
                st[:hi - lo].copy_(self.weights[lo:hi])
            self.state[name] = st
        return st

    def full_state(self, name, barrier=None):
        """The complete [n] state tensor assembled from every rank's slice (checkpointing / membership change).  Staging
        area = the peer-mapped fp32 gradient arena, which must be idle (call between steps); collective."""
        from .peer import _RawBuffer
        st = self.state[name]
        lo, hi = self.slice_range()
        if (lo, hi) == (0, self.n):
            return st[:self.n].clone()
        if self.gdtype is not torch.float32:
            raise RuntimeError("full_state: needs an fp32 gradient arena as staging area")
        world, rank = self.group.world, self.group.rank
        full = torch.empty(self.n, dtype=torch.float32, device=self.device)
        full[lo:hi].copy_(st[:hi - lo])
        self.grads[lo:hi].copy_(st[:hi - lo])
        torch.cuda.synchronize(self.device)
        barrier()
        for r in range(world):
            if r == rank:
                continue
            rlo, rhi = self.slice_range(r, world)
            if rhi <= rlo:
                continue
            src = _RawBuffer(self.g_ptrs[r], self.n * 4, self.device, owner=False).tensor(torch.float32, self.n)
            full[rlo:rhi].copy_(src[rlo:rhi])
        torch.cuda.synchronize(self.device)
        barrier()
        self.grads[lo:hi].zero_()
        return full

    def load_full_state(self, name, full, init_from_weights=False):
        lo, hi = self.slice_range()
        st = self.get_state(name, init_from_weights=init_from_weights)
        if hi > lo:
            st[:hi - lo].copy_(full[lo:hi])

    def bind(self, params):
        for p, v, g in zip(params, self.views, self.grad_views):
            p.data = v
            p.grad = g


def fused_step(arena, mode, lr=0.0, wd=0.0, momentum=0.0, rescale=1.0, clip=None, nesterov=False, multi_precision=True,
               zero_grad=True, beta1=0.9, beta2=0.999, eps=1e-8, bc1=1.0, bc2=1.0, lamb_lower=-1.0, lamb_upper=-1.0,
               one_shot=None, blocks=0, seq=None, hyper=None, nvls=None, ll=None, unroll=None):
    """Launch the fused kernel for one arena on this rank's current stream.

    ``hyper``: fp32 device tensor of HYPER_FLOATS entries (``pack_hyper``) that the kernel reads lr / wd / momentum /
    rescale / clip / bc1 / bc2 and the barrier sequence number from; ``seq`` is then an OFFSET added to the value stored in
    the buffer.  This is what makes a captured launch replayable (parallel/graph.py)."""
    g = arena.group
    lib = g.lib
    _check_abi(lib)
    if _UNROLL[0] is None:
        from ..base import getenv
        _UNROLL[0] = int(getenv("MXB200_AR_UNROLL", 0))
    a = _Args()
    for r in range(g.world):
        a.grads[r] = arena.g_ptrs[r]
        a.weights[r] = arena.w_ptrs[r]
        a.flags[r] = g.flag_ptrs[r]
    a.n = arena.n
    a.rank, a.nranks, a.mode = g.rank, g.world, mode
    if one_shot is not None and bool(one_shot) != arena.one_shot():
        if arena.state and g.world > 1:
            raise RuntimeError("fused_step: the one-shot / two-shot choice fixes the state sharding; do not change it after "
                               "optimizer state exists")
        arena._one_shot = bool(one_shot)
    a.one_shot = int(arena.one_shot())
    a.state_off = arena.slice_range()[0] if not a.one_shot else 0
    use_ll = (ll if ll is not None else True) and g.world > 1 and a.one_shot and arena.gdtype is torch.float32 and \
        arena.n * 4 <= LL_BYTES and mode in (MODE_ALLREDUCE, MODE_SGD, MODE_ADAM) and hasattr(g, "alloc") and \
        getattr(g, "seq_dev", None) is not None and seq is None
    if use_ll:
        if arena.ll_bufs is None:
            # collective allocation: [2 parities][world sources][n] packets of 8 bytes per rank
            arena.ll_bufs = g.alloc(2 * g.world * arena.n * 8)
            arena.ll_ctr = torch.ones(1, dtype=torch.int32, device=arena.device)
        for r in range(g.world):
            a.ll_recv[r] = arena.ll_bufs[1][r]
        a.ll_ctr = arena.ll_ctr.data_ptr()
    # data path: in-switch reduction pays off with many ranks and mid-size messages (measured, tools/bandwidth/sweep.py:
    # N=8: NVLS wins from 256 KB to 256 MB; N=2/4: the unicast peer loads are as fast or faster); nvls=True/False forces it
    gbytes = arena.n * arena.grads.element_size()
    use_nvls = nvls if nvls is not None else (g.world >= 8 and (256 << 10) <= gbytes <= (512 << 20))
    if g.world > 1 and use_nvls and arena.g_mc and mode not in (MODE_LAMB2, MODE_BCAST):
        a.grads_mc = arena.g_mc
    if g.world > 1 and use_nvls and arena.w_mc:
        a.weights_mc = arena.w_mc
    a.g_dt, a.w_dt = code(arena.gdtype), code(arena.wdtype)
    if hyper is not None:
        a.hyper = hyper.data_ptr()
    sd = getattr(g, "seq_dev", None)
    if seq is None and sd is not None:
        # barrier sequence numbers live on the device: the kernel reads and advances them itself (graph-replay safe)
        a.seq_dev = sd.data_ptr()
        a.seq = 0
    elif hyper is not None:
        a.seq = seq if seq is not None else 0
    else:
        a.seq = seq if seq is not None else g.next_seq()
    a.zero_grad = int(zero_grad)
    a.lr, a.wd, a.momentum, a.rescale = lr, wd, momentum, rescale
    a.clip = -1.0 if clip is None else clip
    a.nesterov = int(nesterov)
    a.has_mom = int(momentum != 0.0)
    a.beta1, a.beta2, a.eps, a.bc1, a.bc2 = beta1, beta2, eps, bc1, bc2
    a.lamb_lower, a.lamb_upper = lamb_lower, lamb_upper
    a.chunk_tid = arena.chunk_tid.data_ptr()
    a.pad0 = int(_UNROLL[0] if unroll is None else unroll)
    a.lr_mult = arena.lr_mult.data_ptr()
    a.wd_mult = arena.wd_mult.data_ptr()
    a.err = g.err_dev.value
    need_master = multi_precision and arena.wdtype is not torch.float32
    if mode in (MODE_SGD, MODE_ADAM, MODE_LAMB1, MODE_LAMB2, MODE_BCAST, MODE_LAMB) and need_master:
        a.w32 = arena.get_state("w32", init_from_weights=True).data_ptr()
    if mode == MODE_SGD and momentum != 0.0:
        a.mom = arena.get_state("mom").data_ptr()
    if mode in (MODE_ADAM, MODE_LAMB1):
        a.mom = arena.get_state("mom").data_ptr()
        a.var = arena.get_state("var").data_ptr()
    if mode == MODE_LAMB:
        a.mom = arena.get_state("mom").data_ptr()
        a.var = arena.get_state("var").data_ptr()
        a.rbuf = arena.get_state("rbuf").data_ptr()
        nt = len(arena.layout.names)
        nt_pad = (nt + 255) // 256 * 256
        if arena.norm_bufs is None:
            arena.norm_bufs = g.alloc(2 * nt_pad * 4)            # collective: every rank reaches its first LAMB step together
            arena.norm_tot = torch.zeros(2 * nt_pad, dtype=torch.float32, device=arena.device)
            arena.gsync = torch.zeros(2, dtype=torch.int32, device=arena.device)
        for r in range(g.world):
            a.norms[r] = arena.norm_bufs[1][r]
        a.norm_tot = arena.norm_tot.data_ptr()
        a.gsync = arena.gsync.data_ptr()
        a.nt_pad = nt_pad
    if mode in (MODE_LAMB1, MODE_LAMB2):
        a.rbuf = arena.get_state("rbuf").data_ptr()
        nt = len(arena.layout.names)
        if "norms" not in arena.state:
            arena.state["norms"] = torch.zeros(2 * ((nt + 1023) // 1024 * 1024), dtype=torch.float32, device=arena.device)
        norms = arena.state["norms"]
        a.w_norm2 = norms.data_ptr()
        a.r_norm2 = norms.data_ptr() + norms.numel() // 2 * 4
    with torch.cuda.device(arena.device):
        rc = lib.MXBFusedAllreduce(ctypes.byref(a), c_int(ctypes.sizeof(a)), c_int(blocks), N.stream_of(device=arena.device))
    N.check(lib, rc)
    LAUNCHES[0] += 1
    if zero_grad and mode not in (MODE_LAMB2, MODE_BCAST):
        arena.dirty = False                 # the kernel cleared the gradients it consumed
