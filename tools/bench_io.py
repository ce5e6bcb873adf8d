#!/usr/bin/env python
"""Input-pipeline throughput: images/s of ImageRecordIter on ImageNet-like JPEGs (500x375, quality 90), 224x224 random-
resized-crop + mirror + normalise, for (a) the device pipeline (C++ loader + nvJPEG + fused kernel, batches born on the
GPU as NHWC bf16) and (b) the host pipeline (thread-pool OpenCV decode, fp32 NCHW in pinned memory + H2D copy) -- the
reference's design (src/io/iter_image_recordio_2.cc).  The training step needs ~12.9k img/s per GPU (ResNet-50, bf16).

    python tools/bench_io.py [--images 2048] [--batch 256] [--threads 8] [--out gpurun_out/bench_io.json]
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mxb200 as mx  # noqa: E402
from mxb200 import recordio  # noqa: E402


def make_rec(path, n, seed=0):
    import cv2
    rng = np.random.default_rng(seed)
    w = recordio.MXIndexedRecordIO(path + ".idx", path + ".rec", "w")
    total = 0
    base = [cv2.GaussianBlur(rng.integers(0, 256, (375, 500, 3), dtype=np.uint8), (0, 0), s) for s in (1.0, 2.0, 4.0)]
    for i in range(n):
        img = base[i % 3].copy()
        # texture + structure so the files have a realistic entropy (~60-120 KB)
        x0, y0 = int(rng.integers(0, 400)), int(rng.integers(0, 300))
        cv2.rectangle(img, (x0, y0), (x0 + 90, y0 + 60), [int(v) for v in rng.integers(0, 256, 3)], -1)
        img = np.roll(img, int(rng.integers(0, 500)), axis=1)
        buf = recordio.pack_img(recordio.IRHeader(0, float(i % 1000), i, 0), img, quality=90)
        total += len(buf)
        w.write_idx(i, buf)
    w.close()
    return total / n


def run(it, batches, sync):
    n = 0
    # warm-up epoch part
    for i, b in enumerate(it):
        if i >= 3:
            break
    it.reset()
    sync()
    t0 = time.perf_counter()
    done = 0
    while done < batches:
        for b in it:
            x = b.data[0]
            n += x.shape[0] - (b.pad or 0)
            done += 1
            if done >= batches:
                break
        it.reset()
    sync()
    return n / (time.perf_counter() - t0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=2048)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--batches", type=int, default=24)
    ap.add_argument("--threads", type=int, default=8)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    import torch
    d = tempfile.mkdtemp(prefix="mxb_io_")
    kb = make_rec(os.path.join(d, "train"), args.images) / 1024
    rec, idx = os.path.join(d, "train.rec"), os.path.join(d, "train.idx")
    kw = dict(path_imgrec=rec, path_imgidx=idx, data_shape=(3, 224, 224), batch_size=args.batch, shuffle=True, rand_mirror=True,
              mean_r=123.68, mean_g=116.779, mean_b=103.939, std_r=58.393, std_g=57.12, std_b=57.375,
              preprocess_threads=args.threads, prefetch_buffer=4)
    res = {"images": args.images, "avg_jpeg_kb": round(kb, 1), "batch": args.batch, "threads": args.threads, "cpus": os.cpu_count()}
    if torch.cuda.is_available():
        sync = torch.cuda.synchronize
        it = mx.io.ImageRecordIter(ctx=mx.gpu(0), layout="NHWC", dtype="bfloat16", random_resized_crop=True, min_random_area=0.08,
                                   min_aspect_ratio=0.75, max_aspect_ratio=4.0 / 3, **kw)
        res["device_pipeline_img_s"] = round(run(it, args.batches, sync), 1)
        res["device_pipeline_h2d_bytes_per_image"] = int(it.h2d_bytes / args.batch)
        it.close()
        for be, name in ((0, "default_hybrid"), (3, "hardware")):
            try:
                it = mx.io.gpu_pipeline.GPUImageRecordIter(rec, (3, 224, 224), args.batch, path_imgidx=idx, shuffle=True, rand_mirror=True,
                                                           random_resized_crop=True, min_random_area=0.08, min_aspect_ratio=0.75,
                                                           max_aspect_ratio=4.0 / 3, ctx=mx.gpu(0), layout="NHWC", dtype="bfloat16",
                                                           threads=args.threads, backend=be)
                res["device_pipeline_%s_backend_img_s" % name] = round(run(it, args.batches, sync), 1)
                it.close()
            except Exception as e:  # noqa: BLE001
                res["device_pipeline_%s_backend" % name] = "unavailable: %s" % str(e)[:120]
    else:
        sync = lambda: None  # noqa: E731
    host = mx.io.ImageRecordIter(rand_crop=True, resize=256, **kw)

    class ToDevice(object):      # the reference's consumer copies the fp32 batch to the GPU
        def __init__(self, it):
            self.it = it

        def __iter__(self):
            for b in self.it:
                if torch.cuda.is_available():
                    b.data[0]._t.cuda(non_blocking=True)
                yield b

        def reset(self):
            self.it.reset()
    res["host_pipeline_img_s"] = round(run(ToDevice(host), max(4, args.batches // 4), sync), 1)
    res["host_pipeline_h2d_bytes_per_image"] = 3 * 224 * 224 * 4
    host.close()
    print(json.dumps(res))
    if args.out:
        with open(args.out, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
