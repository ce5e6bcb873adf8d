"""Flat parameter / gradient arenas.

The reference keeps one ``merged`` buffer per key on a greedily chosen GPU (CommDevice::InitMergeBuffer,
src/kvstore/comm.h:678-712) and pushes/pulls ~161 tensors one by one.  Here all tensors of one dtype live
back to back in ONE buffer per role (weights, gradients, fp32 master, optimizer state), each tensor padded
to a 1024-element chunk, with a per-chunk tensor-id table so a single kernel launch can apply per-tensor
lr/wd multipliers.  Model parameters (and their ``.grad``) are *views* into the arena, so autograd writes
gradients straight into the buffer the fused allreduce+update kernel consumes.
"""
import math

import numpy as np
import torch

CHUNK = 1024


class ArenaLayout(object):
    """Placement of a list of tensors (shapes) inside a chunk-padded flat buffer."""

    def __init__(self, names, shapes, align_chunks_to=1):
        self.names = list(names)
        self.shapes = [tuple(s) for s in shapes]
        self.offsets = []
        self.numels = []
        off = 0
        tids = []
        for i, s in enumerate(self.shapes):
            n = int(np.prod(s)) if len(s) else 1
            self.offsets.append(off)
            self.numels.append(n)
            nch = (n + CHUNK - 1) // CHUNK
            tids += [i] * nch
            off += nch * CHUNK
        # pad the chunk count so that it divides evenly among up to `align_chunks_to` ranks
        while len(tids) % max(1, align_chunks_to) != 0:
            tids.append(-1)
            off += CHUNK
        self.total = off
        self.chunk_tid = np.asarray(tids, dtype=np.int32)

    def view(self, flat, i, memory_format_krsc=False):
        s = self.shapes[i]
        v = flat[self.offsets[i]: self.offsets[i] + self.numels[i]]
        if memory_format_krsc and len(s) == 4:
            o, c, kh, kw = s
            return v.view(o, kh, kw, c).permute(0, 3, 1, 2)
        return v.view(s)


class ParamArena(object):
    """Weights + grads (+ optimizer state) of all tensors of ONE dtype on ONE device."""

    def __init__(self, names, tensors, device, alloc=None, align_ranks=8):
        self.device = torch.device(device)
        self.dtype = tensors[0].dtype
        self.layout = ArenaLayout(names, [t.shape for t in tensors], align_chunks_to=align_ranks)
        alloc = alloc or (lambda n, dt: torch.zeros(n, dtype=dt, device=self.device))
        n = self.layout.total
        self.weights = alloc(n, self.dtype)
        self.grads = alloc(n, self.dtype)
        self.chunk_tid = torch.from_numpy(self.layout.chunk_tid).to(self.device)
        self.lr_mult = torch.ones(len(names), dtype=torch.float32, device=self.device)
        self.wd_mult = torch.ones(len(names), dtype=torch.float32, device=self.device)
        self.state = {}
        self.views = []
        self.grad_views = []
        with torch.no_grad():
            for i, t in enumerate(tensors):
                krsc = t.dim() == 4 and not t.is_contiguous() and t.permute(0, 2, 3, 1).is_contiguous()
                v = self.layout.view(self.weights, i, krsc)
                v.copy_(t)
                self.views.append(v)
                self.grad_views.append(self.layout.view(self.grads, i, krsc))

    @property
    def numel(self):
        return self.layout.total

    def get_state(self, name, dtype=torch.float32, init_from_weights=False):
        st = self.state.get(name)
        if st is None:
            st = torch.zeros(self.layout.total, dtype=dtype, device=self.device)
            if init_from_weights:
                st.copy_(self.weights)
            self.state[name] = st
        return st

    def set_multipliers(self, lr_mult, wd_mult):
        self.lr_mult.copy_(torch.as_tensor(lr_mult, dtype=torch.float32))
        self.wd_mult.copy_(torch.as_tensor(wd_mult, dtype=torch.float32))

    def bind(self, params):
        """Point torch Parameters / tensors at the arena views (data and grad)."""
        for p, v, g in zip(params, self.views, self.grad_views):
            p.data = v
            p.grad = g


def group_by_dtype(named_tensors):
    groups = {}
    for name, t in named_tensors:
        groups.setdefault(t.dtype, []).append((name, t))
    return groups
