#!/usr/bin/env python
"""kvstore push+pull bandwidth sweep, 1 KB - 1 GB, fused kernels vs NCCL (BASELINE.json config 5).

Run under torchrun, one rank per GPU:

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/bandwidth/sweep.py --out profiles/bandwidth_sweep_n8.json

For every message size S (fp32 gradient bytes per rank) all ranks run the plain-sum mode of the fused kernel
(csrc/cuda/allreduce.cu kAllreduce: out_r = sum_q grad_q on every rank == kv.push + kv.pull without an updater,
the operation tools/bandwidth/measure.py of the reference times per parameter tensor) in four variants:

    ll_one_shot (<= 256 KB)         {value, flag} packets pushed into every peer's buffer, no barrier (latency path)
    p2p_one_shot / p2p_two_shot     16-byte loads / stores on peer-mapped arenas, flag barriers
    nvls_one_shot / nvls_two_shot   multimem.ld_reduce / multimem.st on the multicast mapping (in-switch reduction)

and ``torch.distributed.all_reduce`` (NCCL) on a tensor of the same size.  Timing: CUDA events on the launching
stream, 5 warm-up + ``--iters`` timed launches, max over ranks.  Reported per row: time (us), algorithm bandwidth S/t,
bus bandwidth S/t * 2(n-1)/n (the formula of the reference tool, tools/bandwidth/measure.py:138) and the fraction of
the NVLink roofline (two-shot moves S(n-1)/n per direction per GPU; 770 GB/s measured peer copy, B200_PROFILING.md).
Correctness of every variant is checked against the NCCL result at each size.
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from mxb200.parallel import fused, peer  # noqa: E402

LINK_GBS = 770.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--min-bytes", type=int, default=1024)
    ap.add_argument("--max-bytes", type=int, default=1 << 30)
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--out", default="")
    ap.add_argument("--no-graph", action="store_true", help="skip the CUDA-graph (device-side latency) measurements")
    ap.add_argument("--unroll", type=int, default=None, help="1: four vectors per thread and iteration in the fused kernel")
    args = ap.parse_args()
    rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", lrank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    group = peer.PeerGroup("process", rank, world, dev, exchange=peer.torch_dist_exchange())
    have_mc = group.multicast_ok()
    rows = []
    ok_all = True

    def timeit(fn, iters):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / iters * 1e3], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    def timeit_graph(fn, reps=20, iters=10):
        """Device-side latency: ``reps`` back-to-back launches captured in ONE CUDA graph, replayed ``iters`` times -- no host
        dispatch between launches (the fused kernel reads and advances its barrier sequence number on the device; NCCL
        collectives are graph-capturable too)."""
        st = torch.cuda.Stream(dev)
        st.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(st):
            for _ in range(3):
                fn()
        torch.cuda.current_stream(dev).wait_stream(st)
        torch.cuda.synchronize()
        dist.barrier()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for _ in range(reps):
                fn()
        g.replay()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / (iters * reps) * 1e3], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        del g
        return t.item()

    size = args.min_bytes
    while size <= args.max_bytes:
        nelem = max(8192, size // 4)          # arena granularity: 1024-element chunks x 8 ranks
        nelem = (nelem + 8191) // 8192 * 8192
        real = size // 4
        t0 = torch.zeros(nelem, dtype=torch.float32, device=dev)
        if have_mc:
            wl, wp, wmc = group.alloc_multicast(nelem * 4)
            gl, gp, gmc = group.alloc_multicast(nelem * 4)
            bufs = {"weights": (wl, wp), "grads": (gl, gp), "weights_mc": wmc, "grads_mc": gmc}
        else:
            bufs = {"weights": group.alloc(nelem * 4), "grads": group.alloc(nelem * 4)}
        ar = fused.SymmetricArena(group, ["t"], [t0], bufs, torch.float32, torch.float32)
        g = torch.Generator(device=dev).manual_seed(100 + rank)
        ar.grads.copy_(torch.randn(nelem, device=dev, generator=g))
        ref = ar.grads.clone()
        dist.all_reduce(ref)
        row = {"bytes": size, "elements": real, "world": world}
        iters = args.iters if size <= (64 << 20) else max(5, args.iters // 3)
        variants = [("p2p_one_shot", False, True, False), ("p2p_two_shot", False, False, False)]
        if nelem * 4 <= fused.LL_BYTES:
            variants.insert(0, ("ll_one_shot", False, True, True))      # barrier-free flag-carrying packets
        if have_mc:
            variants += [("nvls_one_shot", True, True, False), ("nvls_two_shot", True, False, False)]
        for name, nv, one, ll in variants:
            if one and size > (8 << 20):
                continue                      # one-shot moves (n-1) x S per rank: pointless for large messages
            def run(nv=nv, one=one, ll=ll):
                fused.fused_step(ar, fused.MODE_ALLREDUCE, rescale=1.0, one_shot=one, zero_grad=False, nvls=nv, ll=ll, unroll=args.unroll)
            run()
            torch.cuda.synchronize()
            err = ((ar.weights - ref).abs().max() / (ref.abs().max() + 1e-12)).item()
            e = torch.tensor([err], device=dev)
            dist.all_reduce(e, op=dist.ReduceOp.MAX)
            ok = e.item() < 1e-5
            ok_all &= ok
            us = timeit(run, iters)
            row[name + "_us"] = us
            row[name + "_ok"] = ok
            if size <= (4 << 20) and not args.no_graph:
                row[name + "_graph_us"] = timeit_graph(run)
            row[name + "_busbw_GBs"] = size / us / 1e3 * 2 * (world - 1) / world
            if not one:
                row[name + "_frac_nvlink_roofline"] = (size * (world - 1) / world / (LINK_GBS * 1e9)) / (us * 1e-6)
        nt = torch.randn(max(1, real), device=dev)
        us = timeit(lambda: dist.all_reduce(nt), iters)
        row["nccl_us"] = us
        row["nccl_busbw_GBs"] = size / us / 1e3 * 2 * (world - 1) / world
        best = min((row[k], k) for k in row if k.endswith("_us") and not k.endswith("_graph_us") and not k.startswith("nccl"))
        row["best_fused"] = best[1][:-3]
        row["best_fused_vs_nccl"] = us / best[0]
        if size <= (4 << 20) and not args.no_graph:
            row["nccl_graph_us"] = timeit_graph(lambda: dist.all_reduce(nt))
            bg = min((row[k], k) for k in row if k.endswith("_graph_us") and not k.startswith("nccl"))
            row["best_fused_graph"] = bg[1][:-9]
            row["best_fused_graph_vs_nccl_graph"] = row["nccl_graph_us"] / bg[0]
        group.check_error()
        rows.append(row)
        if rank == 0:
            print(json.dumps(row), flush=True)
        del ar
        size *= 4
    if rank == 0 and args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        with open(args.out, "w") as f:
            json.dump({"world": world, "multicast": bool(have_mc), "link_GBs_measured": LINK_GBS, "rows": rows}, f, indent=1)
    dist.barrier()
    group.close()
    dist.destroy_process_group()
    sys.exit(0 if ok_all else 1)


if __name__ == "__main__":
    main()
