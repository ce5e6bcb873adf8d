#!/usr/bin/env python
"""BERT pre-training throughput (config 4 of SURVEY.md: BERT-base, bf16, LAMB through the fused arena trainer).

    python tools/bench_bert.py [--model base|large] [--batch 64] [--seq 128] [--masked 20] [--steps 20] [--warmup 5]
    torchrun --nproc-per-node N tools/bench_bert.py ...          (one rank per GPU, fused P2P allreduce + LAMB)

Synthetic token ids, random-init weights.  Device-timed (CUDA events, max over ranks); prints one JSON line.
Per step: forward + backward of the MLM + NSP loss (fused tcgen05 attention with in-kernel dropout, csrc/cuda/attention.cu),
gradient allreduce and the LAMB update in ONE fused launch per arena (the per-tensor norms are reduced across ranks inside
the kernel, csrc/cuda/allreduce.cu mode kLamb), dropout 0.1 active."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mxb200 import ops  # noqa: E402
from mxb200.models import bert  # noqa: E402
from mxb200.parallel.trainer import DeviceTrainer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="base")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--seq", type=int, default=128)
    ap.add_argument("--masked", type=int, default=20)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    args = ap.parse_args()
    rank, world, lrank = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", lrank)
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    spec = bert.bert_base() if args.model == "base" else bert.bert_large()
    P = bert.init_bert_params(spec, dtype=torch.bfloat16, device=dev, seed=0)
    nparams = sum(v.numel() for v in P.values())
    tr = DeviceTrainer(list(P.items()), dev, rank, world)
    W, G = tr.params(), tr.grad_sinks
    B, S, Mp = args.batch, args.seq, args.masked
    g = torch.Generator(device="cpu").manual_seed(rank)
    ids = torch.randint(0, spec.vocab, (B, S), generator=g).to(dev)
    tt = torch.randint(0, 2, (B, S), generator=g).to(dev)
    am = torch.ones(B, S, dtype=torch.long, device=dev)
    mp = torch.stack([torch.randperm(S, generator=g)[:Mp] for _ in range(B)]).to(dev)
    ml = torch.randint(0, spec.vocab, (B, Mp), generator=g).to(dev)
    nl = torch.randint(0, 2, (B,), generator=g).to(dev)

    def step():
        mlm, nsp = bert.bert_pretrain_forward(spec, W, ids, tt, None, mp, training=True, G=G)
        loss = bert.bert_pretrain_loss(spec, mlm, nsp, ml, nl)
        loss.backward()
        tr.step("lamb", lr=1e-3, wd=0.01, rescale=1.0 / world)
        return loss

    for _ in range(args.warmup):
        loss = step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    l0 = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    tr.check()
    if rank == 0:
        seqs = B * world / ms * 1e3
        # 6 * params * tokens (fwd+bwd matmul flops, embeddings excluded) + attention 12*L*S^2*H per sequence
        dense = nparams - spec.vocab_padded * spec.hidden - spec.max_pos * spec.hidden
        flops = 6.0 * dense * B * S + 12.0 * spec.layers * S * S * spec.hidden * B
        print(json.dumps({"metric": "BERT-%s pretraining throughput" % args.model, "value": seqs, "unit": "sequences/sec",
                          "tokens_per_sec": seqs * S, "n_gpus": world, "ms_per_step": ms, "steps": args.steps,
                          "warmup": args.warmup, "dtype": "bf16", "optimizer": "LAMB (fused arena kernels)",
                          "config": {"batch_per_gpu": B, "seq_len": S, "masked_per_seq": Mp, "params": nparams,
                                     "attention": "fused tcgen05 kernel (csrc/cuda/attention.cu)", "lamb": "one fused launch per arena (in-kernel cross-rank norm reduction)"},
                          "model_tflops_per_gpu": flops / ms / 1e9, "gpu_launches": ops.launch_count() - l0,
                          "final_loss": float(loss.detach()), "data": "synthetic token ids, random-init weights"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
