"""Membership changes for the one-process-per-GPU ``kvstore('device')`` (single-box elasticity).

The multi-host path of the reference changes membership inside ps-lite (scheduler + servers).  Inside one
box the same protocol is run by :class:`ElasticGroup`: a C++ scheduler node (csrc/host/pslite.cc) is the
membership authority (host-file diff, removal before addition, dense re-ranking, audit log), worker processes
meet in ``membership_change_barrier`` at epoch boundaries, and afterwards every surviving process

  1. drains its work (engine.stop(): the reference's closest primitive is Engine Stop/Start,
     threaded_engine_perdevice.cc:70-91),
  2. tears down and rebuilds the peer map (IPC handles of the symmetric arenas, mxb200/parallel/peer.py),
  3. re-shards optimizer state and re-broadcasts weights / averaged aux state to the new owner set,
  4. re-applies the dynamic-mini-batch learning-rate rule (mxb200/lr_scheduler.py).
"""
import os

from ..base import MXNetError


def process_group_membership_barrier(kv, env):
    """kvstore('device') over torch.distributed: membership is fixed by the process group, so the barrier is a
    plain barrier.  Use ``ElasticGroup`` (coordinator-backed) for live add/remove."""
    kv._barrier()
    return kv.num_workers


class ElasticGroup(object):
    """Coordinator-backed worker group for single-box elastic data parallelism."""

    def __init__(self, tag=None):
        from ..kvstore.dist import PsNode, GROUP_ALL, GROUP_WORKER
        self._G_ALL, self._G_WORKER = GROUP_ALL, GROUP_WORKER
        self.node = PsNode("worker", tag=tag)
        self.node.start()
        self.is_new_worker = self.node.is_new_worker
        if not self.is_new_worker:
            self.node.barrier(GROUP_ALL)

    @property
    def rank(self):
        return self.node.info()["rank"]

    @property
    def world(self):
        return self.node.info()["num_workers"]

    def membership_barrier(self, epoch):
        """Returns (old_world, new_world).  Raises SystemExit(0) in a worker that was removed."""
        old = self.world
        self.node.membership_change_barrier({"EPOCH_BEGIN": str(epoch)})
        if self.node.info()["terminated"]:
            raise SystemExit(0)
        self.is_new_worker = False
        return old, self.world

    def barrier(self):
        self.node.barrier(self._G_WORKER)

    def close(self):
        self.node.stop()


def reinit_after_membership_change(trainer, old_world, new_world, lr_scheduler=None, global_batch=None):
    """Steps 1-4 of the module docstring for a gluon.Trainer on the fused path."""
    from .. import engine
    import torch
    eng = engine.get()
    eng.stop()
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    trainer._reset_kvstore()                 # un-fuse: arenas unmapped, parameters back in private memory
    eng.start()
    if lr_scheduler is not None and hasattr(lr_scheduler, "on_membership_change") and global_batch is not None:
        lr_scheduler.on_membership_change(global_batch)
    return new_world
