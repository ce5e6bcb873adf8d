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
            wbytes = fused.SymmetricArena.bytes_needed(tensors, dtype)
            gbytes = fused.SymmetricArena.bytes_needed(tensors, gdtype)
            if getattr(group, "multicast_ok", None) is not None and group.multicast_ok():
                # NVLS: arenas bound to a multicast object -> multimem.ld_reduce / multimem.st in the fused kernel
                wl, wp, wmc = group.alloc_multicast(wbytes)
                gl, gp, gmc = group.alloc_multicast(gbytes)
                bufs = {"weights": (wl, wp), "grads": (gl, gp), "weights_mc": wmc, "grads_mc": gmc}
            else:
                bufs = {"weights": group.alloc(wbytes), "grads": group.alloc(gbytes)}
            ar = fused.SymmetricArena(group, names, tensors, bufs, dtype, gdtype)
            ar.lr_mult.copy_(torch.tensor([lr_mult.get(n, 1.0) for n in names], dtype=torch.float32))
            ar.wd_mult.copy_(torch.tensor([wd_mult.get(n, 1.0) for n in names], dtype=torch.float32))
            self.arenas.append(ar)
            for n, v, g in zip(names, ar.views, ar.grad_views):
                g._mxb_arena = ar                   # ops.functional.sink_touch: 'write' gradient semantics per backward pass
                v.requires_grad_(True)
                self.param_views[n] = v
                self.grad_sinks[n] = g
                if g.dtype == v.dtype:
                    v.grad = g                      # autograd accumulates in place into the arena

                    def _touch(grad, sink=g):       # runs BEFORE the accumulation: 'write' semantics per backward pass
                        from ..ops.functional import sink_touch
                        sink_touch(sink)
                        return grad
                    self._hooks.append(v.register_hook(_touch))
                else:
                    # low-precision parameter, fp32 arena: ops that know about sinks write there directly; for
                    # everything else (library convolutions) the freshly produced gradient is folded in here
                    def _move(p, sink=g):
                        if p.grad is not None:
                            from ..ops.functional import sink_touch
                            sink_touch(sink)
                            sink.add_(p.grad)
                            p.grad = None
                    self._hooks.append(v.register_post_accumulate_grad_hook(_move))
        self.num_update = 0
        self.hyper_host = self.hyper_dev = None
        self._aux = {}
        from ..base import getenv as _ge
        self.fused_lamb = bool(_ge("MXB200_FUSED_LAMB", True))

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
                seq0 = 0
                if getattr(self.group, "seq_dev", None) is None:
                    seq0 = self.group.next_seq()
                    if optimizer == "lamb" and not (hasattr(self.group, "alloc") and self.fused_lamb):
                        self.group.next_seq()              # the second launch of the step uses seq0 + 1
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
            elif hasattr(self.group, "alloc") and self.fused_lamb:
                # ONE launch: reduce + moments + direction + per-tensor norms, in-kernel cross-rank norm reduction over
                # peer memory, trust ratio + update + weight fan-out (no NCCL call, no second kernel)
                fused.fused_step(ar, fused.MODE_LAMB, lr=lr_eff, wd=wd, rescale=rescale, clip=clip, beta1=beta1,
                                 beta2=beta2, eps=eps, bc1=bc1, bc2=bc2, multi_precision=multi_precision,
                                 zero_grad=zero_grad, blocks=blocks, hyper=hy, lamb_lower=lamb_bounds[0],
                                 lamb_upper=lamb_bounds[1])
            else:
                ar.get_state("rbuf")
                fused.fused_step(ar, fused.MODE_LAMB1, lr=lr_eff, wd=wd, rescale=rescale, clip=clip, beta1=beta1,
                                 beta2=beta2, eps=eps, bc1=bc1, bc2=bc2, multi_precision=multi_precision,
                                 zero_grad=zero_grad, blocks=blocks, hyper=hy)
                self._allreduce_norms(ar)
                fused.fused_step(ar, fused.MODE_LAMB2, lr=lr_eff, multi_precision=multi_precision, zero_grad=False,
                                 lamb_lower=lamb_bounds[0], lamb_upper=lamb_bounds[1], blocks=blocks, hyper=hy,
                                 seq=1 if (hy is not None and getattr(self.group, "seq_dev", None) is None) else None)
                ar.state["norms"].zero_()
        if phase == "host":
            self.hyper_dev.copy_(self.hyper_host)          # pageable -> device: staged by the runtime, stream ordered

    def _allreduce_norms(self, ar):
        if self.world == 1:
            return
        import torch.distributed as dist
        # tiny message (2 floats per tensor); per-tensor trust ratios need the global norms
        dist.all_reduce(ar.state["norms"])

    # ------------------------------------------------------------------ state re-sharding (elastic membership, checkpoints)
    _STATE_NAMES = ("w32", "mom", "var")

    def gather_state(self, barrier):
        """Kept for API compatibility: the optimizer state is sharded ZeRO-1 style (each rank allocates and updates only
        its slice); ``export_state`` assembles the full tensors itself, so there is nothing to do ahead of it."""
        return None

    def export_state(self, barrier=None):
        """{num_update, params: name -> dict(weight=tensor, <state>=fp32 tensor ...)} as private, arena-layout-independent
        copies.  Collective when world > 1 (every rank's slice is fetched over peer memory); ``barrier`` is a host-level
        barrier over the current ranks (default: torch.distributed.barrier)."""
        if barrier is None and self.world > 1:
            import torch.distributed as dist
            barrier = dist.barrier
        out = {"num_update": self.num_update, "params": {}}
        for ar in self.arenas:
            fulls = {}
            for sname in self._STATE_NAMES:
                if sname in ar.state:
                    fulls[sname] = ar.full_state(sname, barrier)
            for i, n in enumerate(ar.layout.names):
                d = {"weight": ar.views[i].detach().clone()}
                for sname, full in fulls.items():
                    d[sname] = ar.layout.view(full, i, False).detach().clone()
                out["params"][n] = d
        torch.cuda.synchronize(self.device)
        return out

    def import_state(self, state, keep_full=False):
        """Inverse of export_state on a (new) trainer over the same parameter set (any world size / sharding).
        ``keep_full``: keep the assembled full-size state tensors until ``sync_from_rank0`` has served the other ranks."""
        self.num_update = int(state.get("num_update", 0))
        with torch.no_grad():
            for ar in self.arenas:
                present = set()
                for n in ar.layout.names:
                    present.update(k for k in state["params"].get(n, {}) if k in self._STATE_NAMES)
                fulls = {k: torch.zeros(ar.n, dtype=torch.float32, device=self.device) for k in present}
                if "w32" in fulls:
                    fulls["w32"].copy_(ar.weights)
                for i, n in enumerate(ar.layout.names):
                    d = state["params"].get(n)
                    if d is None:
                        continue
                    ar.views[i].copy_(d["weight"].reshape(ar.views[i].shape).to(self.device))
                    for sname in present:
                        if sname in d:
                            tgt = ar.layout.view(fulls[sname], i, False)
                            tgt.copy_(d[sname].reshape(-1).to(self.device).view_as(tgt))
                for sname, full in fulls.items():
                    ar.load_full_state(sname, full, init_from_weights=(sname == "w32"))
                ar._fulls = fulls if keep_full else None

    def sync_from_rank0(self, barrier):
        """Every rank adopts rank 0's weights and its own SLICE of rank 0's full optimizer state (how a newcomer obtains
        live state; the reference's newcomers pull from the parameter servers, python/mxnet/module/module.py +
        kvstore_dist.h).  Rank 0 must hold the full state (``import_state(..., keep_full=True)``, or world was 1 / the arena
        is one-shot); it stages one state tensor at a time in its peer-mapped fp32 gradient arena."""
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
            if ar.gdtype is not torch.float32:
                continue
            fulls = getattr(ar, "_fulls", None) or {}
            g0 = P_._RawBuffer(ar.g_ptrs[0], ar.n * 4, self.device, owner=False).tensor(torch.float32, ar.n)
            lo, hi = ar.slice_range()
            # which state tensors exist is decided by rank 0 (the optimizer in use): a 3-float mask in its gradient arena
            if self.rank == 0:
                mask = [1.0 if (n in fulls or (n in ar.state and ar.slice_range() == (0, ar.n))) else 0.0
                        for n in self._STATE_NAMES]
                ar.grads[:3].copy_(torch.tensor(mask, device=self.device))
            torch.cuda.synchronize(self.device)
            barrier()
            present = [bool(v > 0.5) for v in g0[:3].tolist()]
            torch.cuda.synchronize(self.device)
            barrier()
            for name, has in zip(self._STATE_NAMES, present):
                if not has:
                    continue
                if self.rank == 0:
                    ar.grads.copy_(fulls[name] if name in fulls else ar.state[name][:ar.n])
                torch.cuda.synchronize(self.device)
                barrier()
                if self.rank != 0 and hi > lo:
                    ar.get_state(name)[:hi - lo].copy_(g0[lo:hi])
                torch.cuda.synchronize(self.device)
                barrier()
            ar.grads.zero_()
            ar._fulls = None
        torch.cuda.synchronize(self.device)
        barrier()

    # ------------------------------------------------------------------ small-tensor collectives on the fused kernel
    def _aux_arena(self, numel):
        """fp32 symmetric arena (gradient side = contributions, weight side = result) for small replicated state such as
        BatchNorm running statistics.  Collective allocation: every rank must ask for the same size."""
        n = (int(numel) + 8191) // 8192 * 8192
        ar = self._aux.get(n)
        if ar is None:
            t0 = torch.zeros(n, dtype=torch.float32, device=self.device)
            bufs = {"weights": self.group.alloc(n * 4), "grads": self.group.alloc(n * 4)}
            ar = self._aux[n] = fused.SymmetricArena(self.group, ["aux"], [t0], bufs, torch.float32, torch.float32)
        return ar

    def _allreduce_small(self, tensors, scale, only_rank0=False):
        if self.world == 1 or not tensors:
            return
        flat = torch.cat([t.detach().reshape(-1).float() for t in tensors])
        ar = self._aux_arena(flat.numel())
        m = flat.numel()
        if only_rank0 and self.rank != 0:
            ar.grads[:m].zero_()
        else:
            ar.grads[:m].copy_(flat)
        # ONE launch: cross-GPU sum over peer memory (its own flag barriers order the ranks, no host barrier needed)
        fused.fused_step(ar, fused.MODE_ALLREDUCE, rescale=scale, zero_grad=True)
        res = ar.weights[:m]
        o = 0
        with torch.no_grad():
            for t in tensors:
                t.copy_(res[o:o + t.numel()].reshape(t.shape).to(t.dtype))
                o += t.numel()

    def average_across_ranks(self, tensors):
        """In-place mean over the ranks (BatchNorm running statistics at an epoch end; the reference averages aux states
        over devices in executor_group.py:415-434 and over workers on the servers)."""
        self._allreduce_small(tensors, 1.0 / self.world)

    def broadcast_from_rank0(self, tensors, barrier=None):
        """Small tensors (BatchNorm running statistics, counters): every rank adopts rank 0's values -- a plain-sum launch of
        the fused kernel in which only rank 0 contributes."""
        self._allreduce_small(tensors, 1.0, only_rank0=True)

    def scale_state(self, name, factor):
        """Multiply an optimizer state tensor (e.g. the momentum buffer) by ``factor``: the momentum correction of an abrupt
        learning-rate change at a membership change (Goyal et al. 2017, remark 2)."""
        if factor == 1.0:
            return
        for ar in self.arenas:
            st = ar.state.get(name)
            if st is not None:
                st.mul_(float(factor))

    def check(self):
        self.group.check_error()

    def close(self):
        self.group.close()
