"""ImageRecordIter on the GPU: native RecordIO loader -> compressed bytes over the host link -> nvJPEG batched decode ->
one fused crop / resize / mirror / normalise / cast kernel that writes the network's input tensor.

Reference pipeline (all on CPU threads, fp32 NCHW shipped to the device afterwards):
``src/io/iter_image_recordio_2.cc`` (chunk reader, ``part_index / num_parts`` sharding, shuffle, OpenCV decode on
``preprocess_threads``), ``src/io/image_aug_default.cc`` (resize, rand_crop, random_resized_crop, rand_mirror),
``src/io/iter_normalize.h`` (mean / std), ``iter_batchloader.h`` (round_batch padding), ``iter_prefetcher.h``.

Here (``csrc/host/recordio.cc`` + ``csrc/cuda/jpeg.cu``):

* C++ worker threads pack the JPEG payloads of a batch into a pinned staging slot, ``depth`` batches ahead;
* the frame sizes are read from the JPEG headers on the host, which is all the augmentation parameters need, so the
  random crop windows are drawn before anything is decoded;
* nvJPEG decodes the whole batch into one device arena; ``crop_resize_kernel`` samples each image's window bilinearly
  (resize-shorter-edge + crop collapse into ONE resampling: cropping a bilinearly resized image is bilinear sampling of
  the source at affinely mapped coordinates) and emits NHWC bf16 (our models) or NCHW fp32 (the generic API);
* everything runs on a side stream one batch ahead of the consumer.

Without a GPU (``ctx`` is a CPU context) the same loader and window logic run with an OpenCV decode and a numpy
implementation of the kernel's formula: that is what the CPU unit tests exercise.
"""
import ctypes
import math

import numpy as np
import torch

from .. import _native as N
from ..base import MXNetError, getenv
from ..context import Context, cpu
from ..ndarray.ndarray import NDArray
from ..ops._dtypes import code as _dt
from .io import DataBatch, DataDesc, DataIter

__all__ = ["GPUImageRecordIter", "RecordFile"]

c_void_p, c_int, c_int64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64

WINDOW = np.dtype([("src_off", "<i8"), ("w", "<i4"), ("h", "<i4"), ("x0", "<f4"), ("y0", "<f4"), ("cw", "<f4"), ("ch", "<f4"),
                   ("mirror", "<i4"), ("pad", "<i4")])
assert WINDOW.itemsize == 40


def _np_ptr(a):
    return c_void_p(a.ctypes.data)


class RecordFile(object):
    """mmap'ed RecordIO file with random access (native counterpart of ``recordio.MXIndexedRecordIO`` for reading)."""

    def __init__(self, path, idx_path=None):
        self.lib = N.host()
        self.h = c_void_p()
        N.check(self.lib, self.lib.MXBRecOpen(str(path).encode(), (str(idx_path) if idx_path else "").encode(),
                                              ctypes.byref(self.h)))
        n = c_int64()
        N.check(self.lib, self.lib.MXBRecCount(self.h, ctypes.byref(n)))
        self.count = n.value

    def __len__(self):
        return self.count

    def info(self, i):
        ln, key, off = c_int64(), c_int64(), c_int64()
        N.check(self.lib, self.lib.MXBRecInfo(self.h, c_int64(i), ctypes.byref(ln), ctypes.byref(key), ctypes.byref(off)))
        return ln.value, key.value, off.value

    def read(self, i):
        """Payload of record ``i`` (framing removed): what ``MXRecordIO.read`` returns."""
        ln = self.info(i)[0]
        buf = ctypes.create_string_buffer(max(1, ln))
        N.check(self.lib, self.lib.MXBRecRead(self.h, c_int64(i), buf, c_int64(ln)))
        return buf.raw[:ln]

    def max_batch_bytes(self, batch):
        v = c_int64()
        N.check(self.lib, self.lib.MXBRecMaxBatchBytes(self.h, c_int(batch), ctypes.byref(v)))
        return v.value

    def close(self):
        if self.h:
            self.lib.MXBRecClose(self.h)
            self.h = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


def crop_resize_reference(img, win, oh, ow):
    """numpy version of ``crop_resize_kernel`` for ONE image (uint8 HWC) -> float32 [oh, ow, 3] (before mean / std)."""
    h, w = img.shape[:2]
    sx, sy = win["cw"] / ow, win["ch"] / oh
    ox = np.arange(ow, dtype=np.float32)
    if win["mirror"]:
        ox = ow - 1 - ox
    fx = np.clip(np.float32(win["x0"]) + (ox + np.float32(0.5)) * np.float32(sx) - np.float32(0.5), 0, w - 1).astype(np.float32)
    fy = np.clip(np.float32(win["y0"]) + (np.arange(oh, dtype=np.float32) + np.float32(0.5)) * np.float32(sy) - np.float32(0.5),
                 0, h - 1).astype(np.float32)
    x0, y0 = fx.astype(np.int64), fy.astype(np.int64)
    x1, y1 = np.minimum(x0 + 1, w - 1), np.minimum(y0 + 1, h - 1)
    ax, ay = (fx - x0)[None, :, None], (fy - y0)[:, None, None]
    f = img.astype(np.float32)
    top = f[y0][:, x0] + (f[y0][:, x1] - f[y0][:, x0]) * ax
    bot = f[y1][:, x0] + (f[y1][:, x1] - f[y1][:, x0]) * ax
    return top + (bot - top) * ay


class GPUImageRecordIter(DataIter):
    """``mx.io.ImageRecordIter`` with device-side decode and augmentation.  Extra arguments over the reference's:
    ``ctx`` (where batches are produced), ``layout`` ("NCHW" like the reference, or "NHWC"), ``dtype`` may be
    "bfloat16".  ``set_partition(num_parts, part_index)`` re-shards at the next ``reset()`` (elastic membership)."""

    def __init__(self, path_imgrec, data_shape, batch_size, path_imgidx=None, label_width=1, shuffle=False, num_parts=1,
                 part_index=0, rand_crop=False, rand_mirror=False, random_resized_crop=False, min_random_area=1.0,
                 max_random_area=1.0, min_aspect_ratio=None, max_aspect_ratio=0.0, mean=(0.0, 0.0, 0.0), std=(1.0, 1.0, 1.0),
                 resize=-1, threads=4, prefetch=4, dtype="float32", layout="NCHW", ctx=None, seed=0, round_batch=True,
                 data_name="data", label_name="softmax_label", backend=None):
        super(GPUImageRecordIter, self).__init__(batch_size)
        self.ctx = ctx if ctx is not None else cpu()
        if not isinstance(self.ctx, Context):
            raise MXNetError("ctx must be a Context")
        self.dev = self.ctx.torch_device
        self._cuda = self.dev.type == "cuda"
        if len(data_shape) != 3 or data_shape[0] != 3:
            raise MXNetError("GPUImageRecordIter: data_shape must be (3, H, W)")
        self.data_shape = tuple(int(v) for v in data_shape)
        self.oh, self.ow = self.data_shape[1], self.data_shape[2]
        self.layout = layout.upper()
        if self.layout not in ("NCHW", "NHWC"):
            raise MXNetError("layout must be NCHW or NHWC")
        self._tdtype = {"float32": torch.float32, "bfloat16": torch.bfloat16}[str(np.dtype(dtype) if dtype != "bfloat16" else dtype)]
        self.dtype = dtype
        self.label_width = int(label_width)
        self.shuffle, self.round_batch = bool(shuffle), bool(round_batch)
        self.rand_crop, self.rand_mirror, self.random_resized_crop = bool(rand_crop), bool(rand_mirror), bool(random_resized_crop)
        self.area = (float(min_random_area), float(max_random_area))
        if min_aspect_ratio is not None:
            self.aspect = (float(min_aspect_ratio), float(max_aspect_ratio))
        else:
            self.aspect = (1.0 - float(max_aspect_ratio), 1.0 + float(max_aspect_ratio))
        self.resize = int(resize) if resize and resize > 0 else -1
        self.mean = np.asarray(mean, np.float32)
        self.std = np.asarray(std, np.float32)
        self.seed = int(seed)
        self.num_parts, self.part_index = int(num_parts), int(part_index)
        self.epoch = 0

        self.lib = N.host()
        self.rec = RecordFile(path_imgrec, path_imgidx)
        if len(self.rec) == 0:
            raise MXNetError("empty RecordIO file %s" % path_imgrec)
        self.depth = max(2, int(prefetch))
        self.slot_bytes = max(4096, self.rec.max_batch_bytes(batch_size))
        from .. import storage as _st              # pooled pinned blocks (plain host memory without a GPU)
        self._slots = [_st.alloc_pinned(self.slot_bytes) for _ in range(self.depth)]
        ptrs = (c_void_p * self.depth)(*[c_void_p(t.data_ptr()) for t in self._slots])
        self.loader = c_void_p()
        N.check(self.lib, self.lib.MXBLoaderCreate(self.rec.h, c_int(batch_size), c_int(self.label_width), c_int(max(1, int(threads))),
                                                   c_int(self.depth), ptrs, c_int64(self.slot_bytes), ctypes.byref(self.loader)))
        shape = (batch_size,) + (self.data_shape if self.layout == "NCHW" else (self.oh, self.ow, 3))
        self.provide_data = [DataDesc(data_name, shape, dtype, layout=self.layout)]
        self.provide_label = [DataDesc(label_name, (batch_size, self.label_width) if self.label_width > 1 else (batch_size,),
                                       "float32")]
        # device side
        self._dec = None
        self._stream = torch.cuda.Stream(self.dev) if self._cuda else None
        self._bufs = [None, None]            # per pipeline slot: arena, windows (pinned + device), output, labels, event
        self._flip = 0
        self._staged = None
        self._outstanding = []               # events of batches whose staging slot has not been released yet
        if self._cuda:
            self.cu = N.cuda()
            ok = c_int(0)
            self.cu.MXBJpegAvailable(ctypes.byref(ok))
            if not ok.value:
                raise MXNetError("GPUImageRecordIter: libnvjpeg could not be loaded")
            # nvJPEG back-end: GPU-hybrid (Huffman decode on the GPU as well) is 3.7x the default hybrid back-end on a B200
            # at batch 256 (7.4 k vs 2.0 k img/s, profiles/bench_io_*); fall back to the default if it cannot be created
            want = getenv("MXB200_NVJPEG_BACKEND", None) if backend is None else backend
            order = [int(want)] if want is not None else [2, 0]
            self._dec = None
            err = None
            for be in order:
                dec = c_void_p()
                with torch.cuda.device(self.dev):
                    rc = self.cu.MXBJpegCreate(c_int(be), c_int(max(1, int(threads))), ctypes.byref(dec))
                if rc == 0:
                    self._dec, self.backend = dec, be
                    break
                err = self.cu.MXBGetLastError()
            if self._dec is None:
                raise MXNetError("nvJPEG decoder could not be created: %s" % (err.decode() if err else "?"))
            self._mean_t = torch.tensor(self.mean)
            self._std_t = torch.tensor(self.std)
        self.h2d_bytes = 0
        self._n = batch_size
        self._off = np.zeros(batch_size + 1, np.int64)
        self._len = np.zeros(batch_size, np.int64)
        self._lab = np.zeros(batch_size * self.label_width, np.float32)
        self._ids = np.zeros(batch_size, np.int64)
        self._w = np.zeros(batch_size, np.int32)
        self._h = np.zeros(batch_size, np.int32)
        self._start_epoch()

    # ---- epoch control --------------------------------------------------------------------------------------
    def set_partition(self, num_parts, part_index):
        """New shard from the next ``reset()`` on (the elastic fit loop calls this when the worker set changed)."""
        self.num_parts, self.part_index = int(num_parts), int(part_index)

    def _start_epoch(self):
        self._drain()
        N.check(self.lib, self.lib.MXBLoaderReset(self.loader, c_int(self.num_parts), c_int(self.part_index), c_int(int(self.shuffle)),
                                                  ctypes.c_uint64(self.seed * 1000003 + self.epoch), c_int(int(self.round_batch))))
        self._rng = np.random.default_rng([self.seed, self.part_index, self.epoch])
        self._staged = None

    def reset(self):
        self.epoch += 1
        self._start_epoch()

    def num_batches(self):
        a, b = c_int64(), c_int64()
        N.check(self.lib, self.lib.MXBLoaderNumBatches(self.loader, ctypes.byref(a), ctypes.byref(b)))
        return a.value

    def _drain(self):
        # consume what is staged / outstanding so the loader's slots are all free again
        if self._staged is not None and self._staged[3] is not None:
            self._staged[3].synchronize()
        for ev in self._outstanding:
            if ev is not None:
                ev.synchronize()
            N.check(self.lib, self.lib.MXBLoaderRelease(self.loader))
        self._outstanding = []

    # ---- augmentation parameters (host, from the JPEG headers only) ------------------------------------------------
    def _windows(self, n, w, h):
        win = np.zeros(n, WINDOW)
        win["w"], win["h"] = w[:n], h[:n]
        wf, hf = w[:n].astype(np.float64), h[:n].astype(np.float64)
        rng = self._rng
        if self.random_resized_crop:
            # Inception-style: area fraction and aspect ratio drawn per image, window placed uniformly; ten attempts
            # then the centre crop at the closest valid aspect (image_aug_default.cc:360-420)
            cw, ch = np.zeros(n), np.zeros(n)
            done = np.zeros(n, bool)
            for _ in range(10):
                area = rng.uniform(self.area[0], self.area[1], n) * wf * hf
                logr = rng.uniform(math.log(max(self.aspect[0], 1e-3)), math.log(max(self.aspect[1], 1e-3)), n)
                ar = np.exp(logr)
                tw, th = np.sqrt(area * ar), np.sqrt(area / ar)
                ok = (~done) & (tw <= wf) & (th <= hf) & (tw >= 1) & (th >= 1)
                cw[ok], ch[ok] = tw[ok], th[ok]
                done |= ok
            rest = ~done
            side = np.minimum(wf, hf)
            cw[rest], ch[rest] = side[rest], side[rest]
            x0 = rng.uniform(0, 1, n) * (wf - cw)
            y0 = rng.uniform(0, 1, n) * (hf - ch)
            x0[rest], y0[rest] = ((wf - cw) / 2)[rest], ((hf - ch) / 2)[rest]
        else:
            # optional resize of the shorter edge, then a data_shape window (random position or centred) of the RESIZED image,
            # expressed in source pixels
            if self.resize > 0:
                nw = np.where(hf > wf, self.resize, np.floor(self.resize * wf / hf))
                nh = np.where(hf > wf, np.floor(self.resize * hf / wf), self.resize)
            else:
                nw, nh = wf.copy(), hf.copy()
            # images smaller than the crop are scaled up to cover it (the reference's data_shape check would abort)
            up = np.maximum(np.maximum(self.ow / nw, self.oh / nh), 1.0)
            nw, nh = nw * up, nh * up
            if self.rand_crop:
                px = np.floor(rng.uniform(0, 1, n) * (nw - self.ow + 1))
                py = np.floor(rng.uniform(0, 1, n) * (nh - self.oh + 1))
            else:
                px, py = np.floor((nw - self.ow) / 2), np.floor((nh - self.oh) / 2)
            sx, sy = wf / nw, hf / nh
            x0, y0, cw, ch = px * sx, py * sy, self.ow * sx, self.oh * sy
        win["x0"], win["y0"], win["cw"], win["ch"] = x0, y0, cw, ch
        if self.rand_mirror:
            win["mirror"] = rng.integers(0, 2, n)
        return win

    # ---- one batch -----------------------------------------------------------------------------------------------
    def _stage(self):
        slot, count, pad = c_int(), c_int(), c_int()
        N.check(self.lib, self.lib.MXBLoaderNext(self.loader, ctypes.byref(slot), _np_ptr(self._off), _np_ptr(self._len), _np_ptr(self._lab),
                                                 _np_ptr(self._ids), ctypes.byref(count), ctypes.byref(pad)))
        n = count.value
        if n == 0:
            return None
        blob = self._slots[slot.value]
        N.check(self.lib, self.lib.MXBJpegDims(c_void_p(blob.data_ptr()), _np_ptr(self._off), _np_ptr(self._len), c_int(n), _np_ptr(self._w),
                                               _np_ptr(self._h)))
        bad = np.nonzero(self._w[:n] == 0)[0]
        decoded = {}
        if len(bad) or not self._cuda:
            import cv2
            for i in (range(n) if not self._cuda else bad):
                raw = blob[int(self._off[i]):int(self._off[i] + self._len[i])].numpy()
                a = cv2.imdecode(raw, cv2.IMREAD_COLOR)
                if a is None:
                    raise MXNetError("record %d of the batch is not a decodable image" % i)
                decoded[int(i)] = np.ascontiguousarray(a[:, :, ::-1])
                self._h[i], self._w[i] = a.shape[:2]
        win = self._windows(n, self._w, self._h)
        sizes = (3 * self._w[:n].astype(np.int64) * self._h[:n].astype(np.int64) + 255) // 256 * 256
        out_off = np.zeros(n + 1, np.int64)
        np.cumsum(sizes, out=out_off[1:])
        win["src_off"] = out_off[:n]
        labels = self._lab[:n * self.label_width].reshape((n, self.label_width) if self.label_width > 1 else (n,)).copy()
        self.h2d_bytes = int(self._off[n]) + win.nbytes + labels.nbytes

        if not self._cuda:
            out = np.empty((n, self.oh, self.ow, 3), np.float32)
            for i in range(n):
                out[i] = (crop_resize_reference(decoded[i], win[i], self.oh, self.ow) - self.mean) / self.std
            x = torch.from_numpy(out if self.layout == "NHWC" else np.ascontiguousarray(out.transpose(0, 3, 1, 2))).to(self._tdtype)
            self._outstanding.append(None)
            return (x, torch.from_numpy(labels), pad.value, None, self._ids[:n].copy())

        k = self._flip
        self._flip ^= 1
        b = self._bufs[k]
        total = int(out_off[n])
        if b is None:
            b = {"arena": torch.empty(0, dtype=torch.uint8, device=self.dev), "ev": torch.cuda.Event(),
                 "win_h": torch.empty(self._n * WINDOW.itemsize, dtype=torch.uint8, pin_memory=True),
                 "win_d": torch.empty(self._n * WINDOW.itemsize, dtype=torch.uint8, device=self.dev),
                 "lab_h": torch.empty((self._n, self.label_width) if self.label_width > 1 else (self._n,), dtype=torch.float32,
                                      pin_memory=True)}
            self._bufs[k] = b
        cur = torch.cuda.current_stream(self.dev)
        self._stream.wait_stream(cur)            # the consumer may still be reading this pipeline slot's previous output
        if b["arena"].numel() < total:
            b["arena"] = torch.empty(int(total * 1.25) + 256, dtype=torch.uint8, device=self.dev)
        shape = (n, 3, self.oh, self.ow) if self.layout == "NCHW" else (n, self.oh, self.ow, 3)
        out = torch.empty(shape, dtype=self._tdtype, device=self.dev)
        b["win_h"][:win.nbytes].copy_(torch.from_numpy(win.view(np.uint8)))
        b["lab_h"][:n].copy_(torch.from_numpy(labels))
        with torch.cuda.device(self.dev), torch.cuda.stream(self._stream):
            st = c_void_p(self._stream.cuda_stream)
            good = np.ones(n, bool)
            good[bad] = False
            if good.all():
                N.check(self.cu, self.cu.MXBJpegDecodeBatch(self._dec, c_void_p(blob.data_ptr()), _np_ptr(self._off), _np_ptr(self._len), c_int(n),
                                                            c_void_p(b["arena"].data_ptr()), _np_ptr(out_off), _np_ptr(self._w), st))
            else:
                idx = np.nonzero(good)[0]
                if len(idx):
                    off = np.ascontiguousarray(self._off[idx])
                    ln = np.ascontiguousarray(self._len[idx])
                    oo = np.ascontiguousarray(out_off[idx])
                    ww = np.ascontiguousarray(self._w[idx])
                    N.check(self.cu, self.cu.MXBJpegDecodeBatch(self._dec, c_void_p(blob.data_ptr()), _np_ptr(off), _np_ptr(ln), c_int(len(idx)),
                                                                c_void_p(b["arena"].data_ptr()), _np_ptr(oo), _np_ptr(ww), st))
                for i in bad:                     # PNG / other payloads: decoded on the host, pixels uploaded
                    px = torch.from_numpy(decoded[int(i)].reshape(-1))
                    b["arena"][int(out_off[i]):int(out_off[i]) + px.numel()].copy_(px, non_blocking=False)
            b["win_d"][:win.nbytes].copy_(b["win_h"][:win.nbytes], non_blocking=True)
            lab_d = b["lab_h"][:n].to(self.dev, non_blocking=True)
            N.check(self.cu, self.cu.MXBImageCropResize(c_void_p(b["arena"].data_ptr()), c_void_p(b["win_d"].data_ptr()), c_int(n),
                                                        c_void_p(out.data_ptr()), c_int(_dt(self._tdtype)), c_int(0 if self.layout == "NHWC" else 1),
                                                        c_int(self.oh), c_int(self.ow), _np_ptr(self.mean), _np_ptr(self.std), st))
            from ..ops.functional import LAUNCHES
            LAUNCHES[0] += 1
            b["ev"].record(self._stream)
        self._outstanding.append(b["ev"])
        return (out, lab_d, pad.value, b["ev"], self._ids[:n].copy())

    def _release_done(self, keep):
        # hand staging slots back to the loader once the device has consumed them (always keep `keep` newest outstanding)
        while len(self._outstanding) > keep:
            ev = self._outstanding.pop(0)
            if ev is not None:
                ev.synchronize()
            N.check(self.lib, self.lib.MXBLoaderRelease(self.loader))

    def next(self):
        if self._staged is None:
            self._staged = self._stage()
        cur = self._staged
        if cur is None:
            self._release_done(0)
            raise StopIteration
        x, y, pad, ev, ids = cur
        if ev is not None:
            torch.cuda.current_stream(self.dev).wait_event(ev)
        self._release_done(1)
        self._staged = self._stage()              # batch i+1 decodes while the caller trains on batch i
        return DataBatch([NDArray(x, self.ctx)], [NDArray(y, self.ctx)], pad=pad, index=ids, provide_data=self.provide_data,
                         provide_label=self.provide_label)

    def close(self):
        if getattr(self, "loader", None):
            try:
                self._drain()
            except Exception:  # noqa: BLE001
                pass
            self.lib.MXBLoaderFree(self.loader)
            self.loader = c_void_p()
        if getattr(self, "_dec", None):
            self.cu.MXBJpegFree(self._dec)
            self._dec = None
        if getattr(self, "rec", None):
            self.rec.close()
        if getattr(self, "_slots", None):
            from .. import storage as _st
            for b in self._slots:
                _st.free_pinned(b)
            self._slots = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass
