#!/usr/bin/env python
"""Multi-GPU correctness + timing of the fused P2P allreduce+update kernel (run under torchrun, one rank per GPU).

For each message size: every rank fills its gradient arena with rank-dependent values, all ranks run ONE fused
step (SGD momentum, multi-precision) and the result is compared with an NCCL all_reduce + the same update done in
plain torch fp32.  Then the fused kernel and the NCCL all_reduce(+torch update) are timed on the device
(CUDA events, max over ranks).  Prints one JSON line per size on rank 0 and writes gpurun_out/multigpu_check.json."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mxb200.parallel.trainer import DeviceTrainer  # noqa: E402


def main():
    rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", lrank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    out = []
    ok_all = True
    for nelem in (4096, 262144, 25_600_000):
        torch.manual_seed(1)   # same weights everywhere
        shapes = [nelem // 2, nelem // 4, nelem - nelem // 2 - nelem // 4]
        named = [("w%d" % i, (torch.randn(s, device=dev) * 0.1).bfloat16()) for i, s in enumerate(shapes)]
        ref_w = [t.float().clone() for _, t in named]
        ref_m = [torch.zeros_like(t) for t in ref_w]
        tr = DeviceTrainer(named, dev, rank, world)
        lr, wd, mom = 0.1, 1e-4, 0.9
        for step in range(3):
            gs = []
            for i, (n, _) in enumerate(named):
                g = torch.full((shapes[i],), 0.01 * (rank + 1) * (step + 1), device=dev) + \
                    torch.arange(shapes[i], device=dev, dtype=torch.float32).remainder_(7) * 1e-3
                tr.grad_sinks[n].copy_(g)
                gs.append(g.clone())
            tr.step("sgd", lr=lr, wd=wd, momentum=mom, rescale=1.0 / world)
            tr.check()
            for i, g in enumerate(gs):
                dist.all_reduce(g)
                g = g / world + wd * ref_w[i]
                ref_m[i] = mom * ref_m[i] - lr * g
                ref_w[i] = ref_w[i] + ref_m[i]
        torch.cuda.synchronize()
        err = 0.0
        for i, (n, _) in enumerate(named):
            got = tr.params()[n].detach().float()
            err = max(err, ((got - ref_w[i]).abs().max() / (ref_w[i].abs().max() + 1e-6)).item())
            # gradients must have been consumed (zeroed) by the kernel
            err = max(err, tr.grad_sinks[n].abs().max().item())
        e = torch.tensor([err], device=dev)
        dist.all_reduce(e, op=dist.ReduceOp.MAX)
        ok = e.item() < 1e-2
        ok_all &= ok

        def timeit(fn, iters=20, warm=5):
            for _ in range(warm):
                fn()
            torch.cuda.synchronize()
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return t.item()

        t_fused = timeit(lambda: tr.step("sgd", lr=lr, wd=wd, momentum=mom, rescale=1.0 / world))
        flat = torch.zeros(nelem, device=dev)
        wflat, mflat = torch.zeros(nelem, device=dev), torch.zeros(nelem, device=dev)
        wb = torch.zeros(nelem, device=dev, dtype=torch.bfloat16)

        def nccl_path():
            dist.all_reduce(flat)
            g = flat * (1.0 / world) + wd * wflat
            mflat.mul_(mom).add_(g, alpha=-lr)
            wflat.add_(mflat)
            wb.copy_(wflat)

        t_nccl = timeit(nccl_path)
        t_ar = timeit(lambda: dist.all_reduce(flat))
        row = {"elements": nelem, "grad_bytes": 4 * nelem, "world": world, "max_rel_err": e.item(), "ok": ok,
               "fused_allreduce_update_ms": t_fused, "nccl_allreduce_plus_torch_update_ms": t_nccl,
               "nccl_allreduce_only_ms": t_ar,
               "fused_algbw_GBs": 4 * nelem / t_fused / 1e6}
        # roofline of the fused step = its NVLink traffic at link bandwidth: every rank pulls (n-1)/n of the fp32
        # gradients from its peers and pushes (n-1)/n of the updated bf16 weights back (two-shot); NVLink 5 gives
        # 900 GB/s per direction per GPU, reads and writes travel in opposite directions
        rd = 4.0 * nelem * (world - 1) / world
        wr = 2.0 * nelem * (world - 1) / world
        row["nvlink_read_bytes_per_rank"], row["nvlink_write_bytes_per_rank"] = rd, wr
        row["roofline_ms_at_900GBs"] = max(rd, wr) / 900e9 * 1e3
        row["fraction_of_nvlink_roofline"] = row["roofline_ms_at_900GBs"] / t_fused
        if rank == 0:
            print(json.dumps(row), flush=True)
        out.append(row)
        tr.close()
        del tr
    if rank == 0:
        os.makedirs("gpurun_out", exist_ok=True)
        with open("gpurun_out/multigpu_check_w%d.json" % world, "w") as f:
            json.dump(out, f, indent=1)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok_all else 1)


if __name__ == "__main__":
    main()
