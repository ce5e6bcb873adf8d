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

    def params(self):
        """name -> leaf tensor view into the arena (use these in the model)."""
        return self.param_views

    def zero_grad(self):
        for ar in self.arenas:
            ar.grads.zero_()

    def step(self, optimizer="sgd", lr=0.01, wd=0.0, momentum=0.0, rescale=1.0, clip=None, multi_precision=True,
             beta1=0.9, beta2=0.999, eps=1e-8, zero_grad=True, blocks=0, bias_correction=True, lamb_bounds=(-1.0, -1.0)):
        """One synchronous data-parallel optimizer step (all ranks must call it)."""
        self.num_update += 1
        t = self.num_update
        for ar in self.arenas:
            if optimizer in ("sgd", "nag"):
                fused.fused_step(ar, fused.MODE_SGD, lr=lr, wd=wd, momentum=momentum, rescale=rescale, clip=clip,
                                 nesterov=(optimizer == "nag"), multi_precision=multi_precision, zero_grad=zero_grad,
                                 blocks=blocks)
            elif optimizer == "adam":
                # bias correction folded into lr (python/mxnet/optimizer/optimizer.py:1079-1082)
                lr_t = lr * (1.0 - beta2 ** t) ** 0.5 / (1.0 - beta1 ** t)
                fused.fused_step(ar, fused.MODE_ADAM, lr=lr_t, wd=wd, rescale=rescale, clip=clip, beta1=beta1,
                                 beta2=beta2, eps=eps, multi_precision=multi_precision, zero_grad=zero_grad,
                                 blocks=blocks)
            elif optimizer == "lamb":
                bc1 = 1.0 - beta1 ** t if bias_correction else 1.0
                bc2 = 1.0 - beta2 ** t if bias_correction else 1.0
                ar.get_state("rbuf")
                fused.fused_step(ar, fused.MODE_LAMB1, lr=lr, wd=wd, rescale=rescale, clip=clip, beta1=beta1,
                                 beta2=beta2, eps=eps, bc1=bc1, bc2=bc2, multi_precision=multi_precision,
                                 zero_grad=zero_grad, blocks=blocks)
                self._allreduce_norms(ar)
                fused.fused_step(ar, fused.MODE_LAMB2, lr=lr, multi_precision=multi_precision, zero_grad=False,
                                 lamb_lower=lamb_bounds[0], lamb_upper=lamb_bounds[1], blocks=blocks)
                ar.state["norms"].zero_()
            else:
                raise ValueError("DeviceTrainer: optimizer %r has no fused kernel" % optimizer)

    def _allreduce_norms(self, ar):
        if self.world == 1:
            return
        import torch.distributed as dist
        # tiny message (2 floats per tensor); per-tensor trust ratios need the global norms
        dist.all_reduce(ar.state["norms"])

    def check(self):
        self.group.check_error()

    def close(self):
        self.group.close()
