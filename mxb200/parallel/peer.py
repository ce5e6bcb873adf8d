"""PeerGroup: symmetric, peer-mapped device buffers + signal pads for the fused collectives.

Reference analogue: CommDevice::EnableP2P (src/kvstore/comm.h:715-755) + the per-key merge buffers, and on
the control side the Postoffice group maps that are rebuilt when membership changes
(3rdparty/ps-lite/src/postoffice.cc:33-119).  Three topologies share one interface:

  * ``process``  -- one process per GPU (torchrun / our launcher).  Buffers are cudaMalloc'ed here and
                    exported with CUDA IPC handles; the handles travel over the rendezvous
                    (torch.distributed object all-gather, or the membership coordinator's KV board).
  * ``local``    -- one process drives N GPUs (the reference's kvstore('device') process model):
                    cudaDeviceEnablePeerAccess + raw pointers.
  * ``loopback`` -- N logical ranks on ONE device (unit tests of the kernels on a 1-GPU box).

``rebuild(new_members)`` tears the mapping down and maps the new peer set: the elastic re-init path.
"""
import ctypes

import torch

from .. import _native as N
from ..base import MXNetError, getenv

c_void_p, c_int64, c_int = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int

MAX_RANKS = 8
MAX_BLOCKS = 256
FLAG_BYTES = MAX_BLOCKS * MAX_RANKS * 4


class _RawBuffer(object):
    """A cudaMalloc allocation viewed as torch tensors through __cuda_array_interface__."""

    def __init__(self, ptr, nbytes, device, owner=True):
        self.ptr, self.nbytes, self.device, self.owner = ptr, nbytes, device, owner

    def tensor(self, dtype, numel=None, offset_bytes=0):
        esz = torch.empty((), dtype=dtype).element_size()
        numel = (self.nbytes - offset_bytes) // esz if numel is None else numel
        typestr = {torch.float32: "<f4", torch.int32: "<i4", torch.uint8: "|u1", torch.float16: "<f2",
                   torch.bfloat16: "<i2", torch.int64: "<i8"}[dtype]

        class _CAI(object):
            pass
        o = _CAI()
        o.__cuda_array_interface__ = {"shape": (numel,), "typestr": typestr, "data": (self.ptr + offset_bytes, False),
                                      "version": 2}
        with torch.cuda.device(self.device):
            t = torch.as_tensor(o, device=self.device)
        if dtype is torch.bfloat16:
            t = t.view(torch.bfloat16)
        t._mxb_keepalive = self
        return t


class PeerGroup(object):
    """Ranks 0..n-1 with symmetric allocations.  ``alloc(nbytes)`` returns (local buffer, [ptr per rank])."""

    def __init__(self, mode, rank, world, device, exchange=None):
        self.mode, self.rank, self.world = mode, rank, world
        self.device = torch.device(device)
        self.exchange = exchange          # callable(list_of_bytes_local) -> list over ranks (process mode)
        self.lib = N.cuda()
        self._allocs = []                 # (local _RawBuffer, [peer ptrs], [opened handles])
        self._pooled = {}                 # device pointer -> storage.Handle of blocks that came from the pool
        self._segments = []               # VMM segments (multicast-capable allocations)
        self._mc_ok = None
        self._job = None
        self._mc_tag = 0
        self.seq = 0
        self.err_host = c_void_p()
        self.err_dev = c_void_p()
        N.check(self.lib, self.lib.MXBHostAllocMapped(c_int64(64), ctypes.byref(self.err_host), ctypes.byref(self.err_dev)))
        self.flags_local, self.flag_ptrs = self.alloc(FLAG_BYTES)
        # device-resident barrier sequence {seq, finished CTAs}: starts at 1 on every rank, advanced by the kernels
        self.seq_dev = torch.tensor([1, 0], dtype=torch.int32, device=self.device) if getenv("MXB200_SEQ_DEV", True) else None

    # -- allocation -------------------------------------------------------------------------------
    def _malloc(self, nbytes, device):
        """Zero-filled device block.  By default it comes from the pooled storage manager (csrc/host/storage.cc): when the
        group is rebuilt after a membership change the arenas of the previous peer set are found in the pool instead of
        paying cudaFree + cudaMalloc (``MXB200_POOLED_ARENAS=0``: plain cudaMalloc)."""
        if getenv("MXB200_POOLED_ARENAS", True):
            from .. import storage as _st
            h = _st.alloc(nbytes, torch.device(device))
            with torch.cuda.device(device):
                h.tensor(torch.uint8, numel=nbytes).zero_()
                torch.cuda.current_stream().synchronize()
            self._pooled[h.ptr] = h
            return h.ptr
        p = c_void_p()
        with torch.cuda.device(device):
            N.check(self.lib, self.lib.MXBDeviceMalloc(c_int64(nbytes), ctypes.byref(p)))
        return p.value

    def _free(self, ptr):
        h = self._pooled.pop(ptr, None)
        if h is not None:
            from .. import storage as _st
            _st.free(h)
        else:
            self.lib.MXBDeviceFree(c_void_p(ptr))

    def alloc(self, nbytes):
        nbytes = (int(nbytes) + 255) // 256 * 256
        with torch.cuda.device(self.device):
            local = _RawBuffer(self._malloc(nbytes, self.device), nbytes, self.device)
            opened = []
            if self.world == 1:
                ptrs = [local.ptr]
            elif self.mode == "process":
                h = ctypes.create_string_buffer(64)
                N.check(self.lib, self.lib.MXBIpcGetHandle(c_void_p(local.ptr), h))
                handles = self.exchange(bytes(h.raw))
                ptrs = []
                for r, hb in enumerate(handles):
                    if r == self.rank:
                        ptrs.append(local.ptr)
                        continue
                    out = c_void_p()
                    N.check(self.lib, self.lib.MXBIpcOpenHandle(ctypes.create_string_buffer(hb, 64), ctypes.byref(out)))
                    ptrs.append(out.value)
                    opened.append(out.value)
            else:
                raise MXNetError("PeerGroup.alloc: use LocalPeerGroups for mode %s" % self.mode)
        torch.cuda.synchronize(self.device)
        self._allocs.append((local, ptrs, opened))
        return local, ptrs

    # -- multicast-capable allocation (cuMem VMM, csrc/cuda/vmm.cu) --------------------------------------
    def multicast_ok(self):
        """True when every GPU of the group can bind memory to a multicast object (NVLS: in-switch reduction)."""
        if self.world == 1 or not getenv("MXB200_NVLS", True) or not hasattr(self.lib, "MXBVmmCreate"):
            return False
        if self._mc_ok is None:
            v = c_int(0)
            ok = self.lib.MXBVmmMulticastSupported(c_int(self.device.index or 0), ctypes.byref(v)) == 0 and v.value == 1
            if self.exchange is not None:
                ok = all(self.exchange(b"1" if ok else b"0")[r] == b"1" for r in range(self.world))
            self._mc_ok = bool(ok)
        return self._mc_ok

    def alloc_multicast(self, nbytes):
        """Symmetric allocation that is ALSO bound to a multicast object: returns (local buffer, [unicast ptr per rank],
        multicast ptr).  Physical memory comes from cuMemCreate; handles travel as file descriptors over unix sockets."""
        if self.mode != "process":
            raise MXNetError("alloc_multicast: process-mode groups only (local groups use alloc_multicast_local)")
        lib = self.lib
        dev = self.device.index or 0
        if self._job is None:
            import os as _os
            self._job = self.exchange(_os.urandom(6).hex().encode())[0].decode()
        self._mc_tag += 1
        seg, padded, lptr = c_void_p(), c_int64(0), c_void_p()
        with torch.cuda.device(self.device):
            N.check(lib, lib.MXBVmmCreate(c_int(dev), c_int(self.rank), c_int(self.world), c_int64(int(nbytes)), c_int(1),
                                          ctypes.byref(seg), ctypes.byref(padded), ctypes.byref(lptr)))
            N.check(lib, lib.MXBVmmExchange(seg, self._job.encode(), c_int(self._mc_tag), c_int(1)))
            self.exchange(b"added")                      # every device has been added to the multicast object
            N.check(lib, lib.MXBVmmBindMulticast(seg))
            self.exchange(b"bound")
            ptrs = (c_void_p * self.world)()
            mc = c_void_p()
            N.check(lib, lib.MXBVmmPointers(seg, ptrs, ctypes.byref(mc)))
        local = _RawBuffer(lptr.value, padded.value, self.device, owner=False)
        self._segments.append(seg)
        return local, [int(ptrs[r]) for r in range(self.world)], mc.value

    def check_error(self):
        v = ctypes.cast(self.err_host, ctypes.POINTER(ctypes.c_int))[0]
        if v != 0:
            raise MXNetError("fused collective timed out waiting for peer rank %d (membership out of sync?)" % (v - 1))

    def next_seq(self):
        self.seq += 1
        return self.seq

    def close(self):
        for local, _ptrs, opened in self._allocs:
            for p in opened:
                self.lib.MXBIpcCloseHandle(c_void_p(p))
            if local.owner:
                self._free(local.ptr)
        self._allocs = []
        for seg in self._segments:
            self.lib.MXBVmmFree(seg)
        self._segments = []


def make_local_groups(devices, loopback=False):
    """Single-process groups: one PeerGroup-like object per logical rank sharing raw pointers.

    devices: list of torch devices (``loopback``: the same device repeated)."""
    world = len(devices)
    lib = N.cuda()
    if not loopback:
        for a in devices:
            for b in devices:
                if a != b:
                    N.check(lib, lib.MXBEnablePeerAccess(c_int(torch.device(a).index), c_int(torch.device(b).index)))
    groups = [_LocalGroup(r, world, devices[r]) for r in range(world)]
    for g in groups:
        g._siblings = groups
    flags = _alloc_symmetric(groups, FLAG_BYTES)
    for g, (loc, ptrs) in zip(groups, flags):
        g.flags_local, g.flag_ptrs = loc, ptrs
    return groups


def _alloc_symmetric(groups, nbytes):
    nbytes = (int(nbytes) + 255) // 256 * 256
    locs = []
    for g in groups:
        p = c_void_p()
        with torch.cuda.device(g.device):
            N.check(g.lib, g.lib.MXBDeviceMalloc(c_int64(nbytes), ctypes.byref(p)))
        locs.append(_RawBuffer(p.value, nbytes, g.device))
    for g in groups:
        torch.cuda.synchronize(g.device)
    ptrs = [l.ptr for l in locs]
    for g, l in zip(groups, locs):
        g._allocs.append((l, ptrs, []))
    return [(l, ptrs) for l in locs]


class _LocalGroup(object):
    mode = "local"

    def __init__(self, rank, world, device):
        self.rank, self.world, self.device = rank, world, torch.device(device)
        self.lib = N.cuda()
        self._allocs = []
        self._siblings = None
        self.seq = 0
        self.err_host = c_void_p()
        self.err_dev = c_void_p()
        N.check(self.lib, self.lib.MXBHostAllocMapped(c_int64(64), ctypes.byref(self.err_host), ctypes.byref(self.err_dev)))
        self.flags_local = self.flag_ptrs = None

    check_error = PeerGroup.check_error
    next_seq = PeerGroup.next_seq

    @staticmethod
    def alloc_all(groups, nbytes):
        return _alloc_symmetric(groups, nbytes)

    def close(self):
        for local, _p, _o in self._allocs:
            self.lib.MXBDeviceFree(c_void_p(local.ptr))
        self._allocs = []


def torch_dist_exchange():
    """Handle exchange over torch.distributed (any backend that supports object collectives)."""
    import torch.distributed as dist

    def ex(local_bytes):
        out = [None] * dist.get_world_size()
        dist.all_gather_object(out, local_bytes)
        return out
    return ex


def pslite_exchange(node):
    """Handle exchange through the membership scheduler (csrc/host/pslite.cc kAllGather): works for a worker set
    that changes at run time, which a fixed torch.distributed process group cannot follow."""

    def ex(local_bytes):
        import struct
        rank = node.info()["rank"]
        parts = node.allgather(struct.pack("<i", rank) + bytes(local_bytes))
        parts = sorted(parts, key=lambda b: struct.unpack("<i", b[:4])[0])
        ranks = [struct.unpack("<i", b[:4])[0] for b in parts]
        if ranks != list(range(len(parts))):
            raise MXNetError("handle exchange: ranks are not dense: %s" % (ranks,))
        return [b[4:] for b in parts]
    return ex
