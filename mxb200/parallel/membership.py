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

    def allgather(self, blob):
        return self.node.allgather(blob)

    def peer_group(self, device):
        """Symmetric-memory group over the CURRENT worker set (CUDA IPC handles exchanged through the scheduler)."""
        from . import peer
        return peer.PeerGroup("process", self.rank, self.world, device, exchange=peer.pslite_exchange(self.node))

    def close(self):
        from ..kvstore.dist import K_STOP_SERVER
        try:
            info = self.node.info()
            if not info["terminated"]:
                self.node.barrier(self._G_WORKER)
                if info["rank"] == 0 and info["num_servers"] > 0:
                    self.node.send_command(K_STOP_SERVER, "")
                self.node.barrier(self._G_ALL)
        finally:
            self.node.stop()


class ElasticDeviceTrainer(object):
    """The fused data-parallel trainer on a worker set that changes at epoch boundaries (single box, one process per
    GPU).  Protocol per boundary (reference call stack: SURVEY.md 3.3; here the "servers" are the GPUs themselves):

      survivors + leavers : gather_state (un-shard the ZeRO-1 optimizer state while everyone is still there)
                            -> export private copies -> unmap the arenas
      everyone            : ElasticGroup.membership_barrier (scheduler diffs the host file, removal before addition,
                            launches newcomers, re-ranks densely; leavers exit inside)
      survivors + newcomers: new PeerGroup over the new rank set -> new arenas -> import state
                            -> every rank adopts rank 0's weights / optimizer state / aux (newcomers start from live
                            state exactly like the reference's newcomers pull from the servers)
    """

    def __init__(self, egroup, named_params, device, aux=None, **trainer_kw):
        import torch
        self.eg = egroup
        self.device = torch.device(device)
        self.kw = trainer_kw
        self.aux = list(aux or [])
        self.tr = None
        self._build([(n, t) for n, t in named_params], None)

    def _build(self, named, state):
        from .trainer import DeviceTrainer
        pg = self.eg.peer_group(self.device)
        self.tr = DeviceTrainer(named, self.device, self.eg.rank, self.eg.world, group=pg, **self.kw)
        if state is not None:
            self.tr.import_state(state, keep_full=True)
        self.tr.sync_from_rank0(self.eg.barrier)
        self.tr.broadcast_from_rank0(self.aux, self.eg.barrier)

    def params(self):
        return self.tr.params()

    @property
    def grad_sinks(self):
        return self.tr.grad_sinks

    @property
    def world(self):
        return self.eg.world

    @property
    def rank(self):
        return self.eg.rank

    def step(self, *a, **kw):
        return self.tr.step(*a, **kw)

    def epoch_boundary(self, epoch):
        """Call at the start of every epoch except a newcomer's first.  Returns (old_world, new_world)."""
        import torch
        state = self.tr.export_state(self.eg.barrier)     # full (un-sharded) state on every current member
        named = [(n, d["weight"]) for n, d in state["params"].items()]
        torch.cuda.synchronize(self.device)
        self.eg.barrier()
        self.tr.close()
        old, new = self.eg.membership_barrier(epoch)       # leavers raise SystemExit here
        self._build(named, state)
        return old, new

    def close(self):
        if self.tr is not None:
            self.tr.close()
            self.tr = None


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
