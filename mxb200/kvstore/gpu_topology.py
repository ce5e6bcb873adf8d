"""Topology-aware reduce/broadcast trees (reference: src/kvstore/gpu_topology.h + comm_tree.h, enabled by
MXNET_KVSTORE_USETREE=1).  The tree construction is native (csrc/host/topology.cc); this module queries the link
matrix, turns edges into level schedules and runs tree reduce / broadcast on per-device tensors."""
import ctypes

import numpy as np
import torch

from .. import _native as N
from ..base import MXNetError

ALPHA = 0.7   # link-usage penalty (gpu_topology.h: kLinkUsagePenalty... 0.7)


def link_matrix(device_ids):
    """n x n weights: 0 no peer access, 1 peer access over PCIe, 2+ NVLink (more links -> larger)."""
    n = len(device_ids)
    W = np.zeros((n, n), dtype=np.float32)
    if not torch.cuda.is_available():
        return W + (1 - np.eye(n, dtype=np.float32))
    nvl = None
    try:
        import pynvml
        pynvml.nvmlInit()
        nvl = pynvml
    except Exception:  # noqa: BLE001
        nvl = None
    for i, a in enumerate(device_ids):
        for j, b in enumerate(device_ids):
            if i == j:
                continue
            w = 0.0
            try:
                if torch.cuda.can_device_access_peer(a, b):
                    w = 1.0
            except Exception:  # noqa: BLE001
                w = 0.0
            if w > 0 and nvl is not None:
                try:
                    ha, hb = nvl.nvmlDeviceGetHandleByIndex(a), nvl.nvmlDeviceGetHandleByIndex(b)
                    st = nvl.nvmlDeviceGetP2PStatus(ha, hb, nvl.NVML_P2P_CAPS_INDEX_NVLINK)
                    if st == nvl.NVML_P2P_STATUS_OK:
                        w = 2.0
                except Exception:  # noqa: BLE001
                    pass
            W[i, j] = w
    return W


def compute_trees(W, alpha=ALPHA):
    """-> list over roots of level schedules: trees[root] = [[(parent, child), ...] for level 0, 1, ...]."""
    W = np.ascontiguousarray(W, dtype=np.float32)
    n = W.shape[0]
    if W.shape != (n, n):
        raise MXNetError("link matrix must be square")
    if n == 1:
        return [[]]
    lib = N.host()
    out = np.zeros((n, n - 1, 3), dtype=np.int32)
    N.check(lib, lib.MXBTopoComputeTrees(ctypes.c_int(n), W.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
                                         ctypes.c_float(alpha), out.ctypes.data_as(ctypes.POINTER(ctypes.c_int))))
    trees = []
    for r in range(n):
        levels = {}
        for lvl, p, c in out[r]:
            levels.setdefault(int(lvl), []).append((int(p), int(c)))
        trees.append([levels[k] for k in sorted(levels)])
    return trees


class TreeComm(object):
    """Reduce to / broadcast from a root device along the precomputed tree (CommDeviceTree's role)."""

    def __init__(self, device_ids, W=None):
        self.device_ids = list(device_ids)
        self.W = link_matrix(self.device_ids) if W is None else np.asarray(W, dtype=np.float32)
        self.trees = compute_trees(self.W)

    def reduce(self, tensors, root=0):
        """tensors[i] lives on device i; returns the sum on device ``root`` (inputs are not modified)."""
        bufs = [None] * len(tensors)

        def buf(i):
            if bufs[i] is None:
                bufs[i] = tensors[i].clone()
            return bufs[i]
        for level in reversed(self.trees[root]):      # deepest level first
            for p, c in level:
                src = bufs[c] if bufs[c] is not None else tensors[c]
                buf(p).add_(src.to(buf(p).device, non_blocking=True))
        return buf(root)

    def broadcast(self, value, outs, root=0):
        """value on device root -> every outs[i]."""
        have = {root: value}
        outs[root].copy_(value)
        for level in self.trees[root]:
            for p, c in level:
                outs[c].copy_(have[p], non_blocking=True)
                have[c] = outs[c]
