"""DeviceTrainer: the data-parallel update engine behind ``kvstore('device')`` + ``gluon.Trainer`` /
``Module.update`` when parameters live in symmetric arenas.

It owns, per dtype group, a :class:`SymmetricArena` (weights + grads peer-mapped across ranks, fp32
master / momentum local) and turns one optimizer step into one fused kernel launch per group
(reduce over NVLink + rescale/clip + update + weight fan-out; see csrc/cuda/allreduce.cu).

Reference call path it replaces: gluon/trainer.py:304-374 (_allreduce_grads + _update) and
model.py:183-205 (_update_params_on_kvstore), i.e. per-parameter push/pull + per-parameter update ops.
"""
import torch

from . import fused, peer
from .arena import group_by_dtype

_MODES = {"sgd": fused.MODE_SGD, "nag": fused.MODE_SGD, "adam": fused.MODE_ADAM, "lamb": fused.MODE_LAMB1}


class DeviceTrainer(object):
    def __init__(self, named_params, device, rank=0, world=1, group=None, lr_mult=None, wd_mult=None,
                 grad_dtype=None):
        """named_params: list of (name, tensor) in a deterministic order identical on every rank."""
        self.device = torch.device(device)
        self.rank, self.world = rank, world
        if group is None:
            ex = peer.torch_dist_exchange() if world > 1 else None
            group = peer.PeerGroup("process", rank, world, self.device, exchange=ex)
        self.group = group
        self.arenas = []
        self.param_views = {}
        self.names = [n for n, _ in named_params]
        lr_mult = lr_mult or {}
        wd_mult = wd_mult or {}
        self.grad_sinks = {}
        self._hooks = []
        for dtype, items in group_by_dtype(named_params).items():
            names = [n for n, _ in items]
            tensors = [t for _, t in items]
            # gradients are always accumulated in fp32 (better numerics than bf16 accumulation, and it lets the
            # wgrad GEMM / BN backward kernels reduce straight into the arena); weights keep their own dtype
            gdtype = grad_dtype or (torch.float32 if dtype in (torch.bfloat16, torch.float16) else dtype)
            bufs = {"weights": group.alloc(fused.SymmetricArena.bytes_needed(tensors, dtype)),
                    "grads": group.alloc(fused.SymmetricArena.bytes_needed(tensors, gdtype))}
            ar = fused.SymmetricArena(group, names, tensors, bufs, dtype, gdtype)
            ar.lr_mult.copy_(torch.tensor([lr_mult.get(n, 1.0) for n in names], dtype=torch.float32))
            ar.wd_mult.copy_(torch.tensor([wd_mult.get(n, 1.0) for n in names], dtype=torch.float32))
            self.arenas.append(ar)
            for n, v, g in zip(names, ar.views, ar.grad_views):
                v.requires_grad_(True)
                self.param_views[n] = v
                self.grad_sinks[n] = g
                if g.dtype == v.dtype:
                    v.grad = g                      # autograd accumulates in place into the arena
                else:
                    # low-precision parameter, fp32 arena: ops that know about sinks write there directly; for
                    # everything else (library convolutions) the freshly produced gradient is folded in here
                    def _move(p, sink=g):
                        if p.grad is not None:
                            sink.add_(p.grad)
                            p.grad = None
                    self._hooks.append(v.register_post_accumulate_grad_hook(_move))
        self.num_update = 0
        self.hyper_host = self.hyper_dev = None

    def params(self):
        """name -> leaf tensor view into the arena (use these in the model)."""
        return self.param_views

    def refresh_master(self):
        """The low-precision weights were overwritten from outside (set_params / load): drop the fp32 master copies so
        the next step re-derives them from the weights instead of resurrecting the old values."""
        for ar in self.arenas:
            ar.state.pop("w32", None)

    def zero_grad(self):
        for ar in self.arenas:
            ar.grads.zero_()

    def step(self, optimizer="sgd", lr=0.01, wd=0.0, momentum=0.0, rescale=1.0, clip=None, multi_precision=True,
             beta1=0.9, beta2=0.999, eps=1e-8, zero_grad=True, blocks=0, bias_correction=True, lamb_bounds=(-1.0, -1.0),
             phase=None):
        """One synchronous data-parallel optimizer step (all ranks must call it).

        ``phase`` splits the step for CUDA-graph replay (parallel/graph.py):
          None      eager: scalars travel by value with the launch;
          "host"    bookkeeping only: bump the update counter, reserve barrier sequence numbers and upload this step's
                    scalars (lr, wd, momentum, rescale, clip, bias corrections, seq) into the device-resident hyper buffer;
          "device"  launches only, reading those scalars from the hyper buffer -- the part that is captured / replayed."""
        if phase != "device":
            self.num_update += 1
        t = self.num_update
        if optimizer not in ("sgd", "nag", "adam", "lamb"):
            raise ValueError("DeviceTrainer: optimizer %r has no fused kernel" % optimizer)
        if phase is not None and self.hyper_dev is None:
            self.hyper_host = torch.zeros((len(self.arenas), fused.HYPER_FLOATS), dtype=torch.float32)
            self.hyper_dev = torch.zeros((len(self.arenas), fused.HYPER_FLOATS), dtype=torch.float32, device=self.device)
        for idx, ar in enumerate(self.arenas):
            lr_eff, bc1, bc2 = lr, 1.0, 1.0
            if optimizer == "adam":
                # bias correction folded into lr (python/mxnet/optimizer/optimizer.py:1079-1082)
                lr_eff = lr * (1.0 - beta2 ** t) ** 0.5 / (1.0 - beta1 ** t)
            elif optimizer == "lamb":
                bc1 = 1.0 - beta1 ** t if bias_correction else 1.0
                bc2 = 1.0 - beta2 ** t if bias_correction else 1.0
            if phase == "host":
                seq0 = self.group.next_seq()
                if optimizer == "lamb":
                    self.group.next_seq()                  # the second launch of the step uses seq0 + 1
                self.hyper_host[idx] = torch.from_numpy(fused.pack_hyper(lr_eff, wd, momentum, rescale, clip, seq0, bc1, bc2))
                continue
            hy = self.hyper_dev[idx] if phase == "device" else None
            if optimizer in ("sgd", "nag"):
                fused.fused_step(ar, fused.MODE_SGD, lr=lr_eff, wd=wd, momentum=momentum, rescale=rescale, clip=clip,
                                 nesterov=(optimizer == "nag"), multi_precision=multi_precision, zero_grad=zero_grad,
                                 blocks=blocks, hyper=hy)
            elif optimizer == "adam":
                fused.fused_step(ar, fused.MODE_ADAM, lr=lr_eff, wd=wd, rescale=rescale, clip=clip, beta1=beta1,
                                 beta2=beta2, eps=eps, multi_precision=multi_precision, zero_grad=zero_grad,
                                 blocks=blocks, hyper=hy)
            else:
                ar.get_state("rbuf")
                fused.fused_step(ar, fused.MODE_LAMB1, lr=lr_eff, wd=wd, rescale=rescale, clip=clip, beta1=beta1,
                                 beta2=beta2, eps=eps, bc1=bc1, bc2=bc2, multi_precision=multi_precision,
                                 zero_grad=zero_grad, blocks=blocks, hyper=hy)
                self._allreduce_norms(ar)
                fused.fused_step(ar, fused.MODE_LAMB2, lr=lr_eff, multi_precision=multi_precision, zero_grad=False,
                                 lamb_lower=lamb_bounds[0], lamb_upper=lamb_bounds[1], blocks=blocks, hyper=hy,
                                 seq=1 if hy is not None else None)
                ar.state["norms"].zero_()
        if phase == "host":
            self.hyper_dev.copy_(self.hyper_host)          # pageable -> device: staged by the runtime, stream ordered

    def _allreduce_norms(self, ar):
        if self.world == 1:
            return
        import torch.distributed as dist
        # tiny message (2 floats per tensor); per-tensor trust ratios need the global norms
        dist.all_reduce(ar.state["norms"])

    # ------------------------------------------------------------------ state re-sharding (elastic membership)
    _STATE_NAMES = ("w32", "mom", "var")

    def _slice(self, ar, rank, world):
        """Element range of the arena owned by ``rank`` in the two-shot kernel (csrc/cuda/allreduce.cu: per =
        ceil(n_chunks / world) chunks of 1024 elements); one-shot arenas are replicated."""
        from . import fused as F_
        gbytes = ar.n * ar.grads.element_size()
        if world == 1 or gbytes <= F_.ONE_SHOT_BYTES:
            return 0, ar.n
        n_chunks = ar.n // 1024
        per = (n_chunks + world - 1) // world
        lo = min(n_chunks, per * rank)
        return lo * 1024, min(n_chunks, lo + per) * 1024

    def gather_state(self, barrier):
        """Make every rank hold the FULL optimizer state (fp32 master weights, momentum, second moment), which the
        fused kernel keeps sharded ZeRO-1 style.  Staging area = the peer-mapped fp32 gradient arena (must be idle:
        call between steps); ``barrier`` is a host-level barrier over the current ranks.  Needed before a
        membership change: ranks that leave take their shard with them."""
        from . import peer as P_
        if self.world == 1:
            return
        for ar in self.arenas:
            if ar.gdtype is not torch.float32:
                continue
            lo, hi = self._slice(ar, self.rank, self.world)
            if (lo, hi) == (0, ar.n):
                continue                       # replicated (one-shot) arena: nothing is sharded
            for name in self._STATE_NAMES:
                st = ar.state.get(name)
                if st is None:
                    continue
                ar.grads[lo:hi].copy_(st[lo:hi])
                torch.cuda.synchronize(self.device)
                barrier()
                for r in range(self.world):
                    if r == self.rank:
                        continue
                    rlo, rhi = self._slice(ar, r, self.world)
                    if rhi <= rlo:
                        continue
                    src = P_._RawBuffer(ar.g_ptrs[r], ar.n * 4, self.device, owner=False).tensor(torch.float32, ar.n)
                    st[rlo:rhi].copy_(src[rlo:rhi])
                torch.cuda.synchronize(self.device)
                barrier()
            ar.grads.zero_()
        torch.cuda.synchronize(self.device)

    def export_state(self):
        """name -> dict(weight=tensor, <state>=fp32 tensor ...) as private copies (call after gather_state)."""
        out = {"num_update": self.num_update, "params": {}}
        for ar in self.arenas:
            for i, n in enumerate(ar.layout.names):
                d = {"weight": ar.views[i].detach().clone()}
                for sname in self._STATE_NAMES:
                    st = ar.state.get(sname)
                    if st is not None:
                        d[sname] = ar.layout.view(st, i, False).detach().clone()
                out["params"][n] = d
        return out

    def import_state(self, state):
        """Inverse of export_state on a (new) trainer over the same parameter set."""
        self.num_update = int(state.get("num_update", 0))
        with torch.no_grad():
            for ar in self.arenas:
                for i, n in enumerate(ar.layout.names):
                    d = state["params"].get(n)
                    if d is None:
                        continue
                    ar.views[i].copy_(d["weight"].reshape(ar.views[i].shape))
                    for sname in self._STATE_NAMES:
                        if sname in d:
                            full = ar.get_state(sname, init_from_weights=(sname == "w32"))
                            ar.layout.view(full, i, False).copy_(d[sname].reshape(-1).view_as(ar.layout.view(full, i, False)))

    def sync_from_rank0(self, barrier):
        """Every rank adopts rank 0's weights and optimizer state (how a newcomer obtains live state; the reference's
        newcomers pull from the parameter servers, python/mxnet/module/module.py + kvstore_dist.h).  Uses peer reads of
        rank 0's arenas: weights directly, state staged through its fp32 gradient arena."""
        from . import peer as P_
        if self.world == 1:
            return
        for ar in self.arenas:
            esz = ar.weights.element_size()
            torch.cuda.synchronize(self.device)
            barrier()
            if self.rank != 0:
                src = P_._RawBuffer(ar.w_ptrs[0], ar.n * esz, self.device, owner=False).tensor(ar.wdtype, ar.n)
                ar.weights.copy_(src)
            # which state tensors exist is decided by rank 0 (the optimizer in use); all ranks walk the same list
            for name in self._STATE_NAMES:
                has = torch.tensor([1 if name in ar.state else 0], device=self.device, dtype=torch.float32)
                if ar.gdtype is not torch.float32:
                    continue
                if self.rank == 0:
                    ar.grads[:1].copy_(has)
                torch.cuda.synchronize(self.device)
                barrier()
                g0 = P_._RawBuffer(ar.g_ptrs[0], ar.n * 4, self.device, owner=False).tensor(torch.float32, ar.n)
                present = bool(g0[:1].item() > 0.5)
                torch.cuda.synchronize(self.device)
                barrier()
                if not present:
                    continue
                if self.rank == 0:
                    ar.grads.copy_(ar.state[name])
                torch.cuda.synchronize(self.device)
                barrier()
                if self.rank != 0:
                    ar.get_state(name).copy_(g0)
                torch.cuda.synchronize(self.device)
                barrier()
            ar.grads.zero_()
        torch.cuda.synchronize(self.device)
        barrier()

    def broadcast_from_rank0(self, tensors, barrier):
        """Small fp32/any-dtype tensors (BatchNorm running statistics, counters): every rank adopts rank 0's values.
        Staged through the first fp32 gradient arena in chunks."""
        from . import peer as P_
        if self.world == 1 or not tensors:
            return
        ar = next(a for a in self.arenas if a.gdtype is torch.float32)
        g0 = P_._RawBuffer(ar.g_ptrs[0], ar.n * 4, self.device, owner=False).tensor(torch.float32, ar.n)
        flat = torch.cat([t.detach().reshape(-1).float() for t in tensors])
        for off in range(0, flat.numel(), ar.n):
            m = min(ar.n, flat.numel() - off)
            if self.rank == 0:
                ar.grads[:m].copy_(flat[off:off + m])
            torch.cuda.synchronize(self.device)
            barrier()
            if self.rank != 0:
                flat[off:off + m].copy_(g0[:m])
            torch.cuda.synchronize(self.device)
            barrier()
        ar.grads.zero_()
        o = 0
        with torch.no_grad():
            for t in tensors:
                t.copy_(flat[o:o + t.numel()].reshape(t.shape).to(t.dtype))
                o += t.numel()

    def check(self):
        self.group.check_error()

    def close(self):
        self.group.close()
