"""Host-memory ImageRecordIter: native mmap RecordIO reader (csrc/host/recordio.cc) -> thread-pool JPEG decode + augment
-> batch assembly -> prefetch queue (reference pipeline: src/io/iter_image_recordio_2.cc, image_aug_default.cc,
iter_batchloader.h, iter_prefetcher.h).  Decode runs in OpenCV (GIL released), batches are assembled into pinned host
memory so the device copy can be asynchronous.  The device-side pipeline (nvJPEG + fused augmentation kernel) is
io/gpu_pipeline.py; ``mx.io.ImageRecordIter(ctx=mx.gpu(i))`` selects it."""
import concurrent.futures as cf
import queue
import random
import threading

import numpy as np
import torch

from .. import recordio
from ..ndarray import NDArray, array
from .io import DataIter, DataBatch, DataDesc


class ImageRecordIterImpl(DataIter):
    def __init__(self, path_imgrec, data_shape, batch_size, path_imgidx=None, label_width=1, shuffle=False, num_parts=1,
                 part_index=0, rand_crop=False, rand_mirror=False, mean=(0, 0, 0), std=(1, 1, 1), resize=0, threads=4,
                 prefetch=4, dtype="float32", data_name="data", label_name="softmax_label"):
        super(ImageRecordIterImpl, self).__init__(batch_size)
        self.data_shape = tuple(data_shape)
        self.label_width = label_width
        self.shuffle, self.rand_crop, self.rand_mirror, self.resize = shuffle, rand_crop, rand_mirror, resize
        self.mean = np.asarray(mean, np.float32).reshape(3, 1, 1)
        self.std = np.asarray(std, np.float32).reshape(3, 1, 1)
        self.dtype = dtype
        try:
            from .gpu_pipeline import RecordFile
            self._rf = RecordFile(path_imgrec, path_imgidx)
            self.records = list(range(len(self._rf)))
        except Exception:  # noqa: BLE001  (host library unavailable: python scan)
            self._rf = None
            self.records = self._index(path_imgrec, path_imgidx)
        n = len(self.records) // num_parts
        self.records = self.records[part_index * n:(part_index + 1) * n]
        self.path = path_imgrec
        self.pool = cf.ThreadPoolExecutor(max(1, threads))
        self.prefetch = max(1, prefetch)
        self.provide_data = [DataDesc(data_name, (batch_size,) + self.data_shape, dtype)]
        self.provide_label = [DataDesc(label_name, (batch_size, label_width) if label_width > 1 else (batch_size,), "float32")]
        self._q = None
        self._thread = None
        self._stop = threading.Event()
        self.reset()

    @staticmethod
    def _index(path, idx_path):
        """Offsets of every record (from the .idx file when present, else one sequential scan)."""
        if idx_path:
            offs = []
            with open(idx_path) as f:
                for line in f:
                    offs.append(int(line.strip().split("\t")[1]))
            return offs
        import struct
        offs = []
        with open(path, "rb") as f:
            pos = 0
            while True:
                head = f.read(8)
                if len(head) < 8:
                    break
                _m, lrec = struct.unpack("<II", head)
                cflag, length = lrec >> 29, lrec & ((1 << 29) - 1)
                if cflag in (0, 1):
                    offs.append(pos)
                skip = length + (4 - length % 4) % 4
                f.seek(skip, 1)
                pos += 8 + skip
        return offs

    def _decode(self, raw):
        import cv2
        header, img = recordio.unpack(raw)
        a = cv2.imdecode(np.frombuffer(img, np.uint8), cv2.IMREAD_COLOR)[:, :, ::-1]
        c, h, w = self.data_shape
        if self.resize > 0:
            s = self.resize / min(a.shape[:2])
            a = cv2.resize(a, (max(w, int(round(a.shape[1] * s))), max(h, int(round(a.shape[0] * s)))))
        if a.shape[0] < h or a.shape[1] < w:
            a = cv2.resize(a, (max(w, a.shape[1]), max(h, a.shape[0])))
        if self.rand_crop:
            y0, x0 = random.randint(0, a.shape[0] - h), random.randint(0, a.shape[1] - w)
        else:
            y0, x0 = (a.shape[0] - h) // 2, (a.shape[1] - w) // 2
        a = a[y0:y0 + h, x0:x0 + w]
        if self.rand_mirror and random.random() < 0.5:
            a = a[:, ::-1]
        a = (a.transpose(2, 0, 1).astype(np.float32) - self.mean) / self.std
        return a, header.label

    def _producer(self, order):
        f = recordio.MXRecordIO(self.path, "r") if self._rf is None else None
        bs = self.batch_size
        try:
            for b0 in range(0, len(order), bs):
                if self._stop.is_set():
                    return
                idx = order[b0:b0 + bs]
                raws = []
                for off in idx:
                    if f is None:
                        raws.append(self._rf.read(off))
                        continue
                    f.handle.seek(off)
                    raws.append(f.read())
                outs = list(self.pool.map(self._decode, raws))
                pad = bs - len(outs)
                data = torch.empty((bs,) + self.data_shape, dtype=torch.float32, pin_memory=torch.cuda.is_available())
                label = torch.zeros((bs, self.label_width) if self.label_width > 1 else (bs,), dtype=torch.float32)
                for i, (a, l) in enumerate(outs):
                    data[i] = torch.from_numpy(np.ascontiguousarray(a))
                    label[i] = torch.as_tensor(l)
                for i in range(len(outs), bs):
                    data[i] = data[i - len(outs)]
                    label[i] = label[i - len(outs)]
                self._q.put(DataBatch([NDArray(data)], [NDArray(label)], pad=pad, provide_data=self.provide_data,
                                      provide_label=self.provide_label))
        finally:
            if f is not None:
                f.close()
            self._q.put(None)

    def _stop_producer(self):
        if self._thread is not None:
            self._stop.set()
            while self._thread.is_alive():
                try:
                    self._q.get(timeout=0.05)
                except queue.Empty:
                    pass
            self._thread.join()
            self._thread = None

    def close(self):
        """Stop the producer thread (a daemon thread killed inside a native call at interpreter exit aborts the process)."""
        self._stop_producer()
        self.pool.shutdown(wait=True)

    def __del__(self):
        try:
            self._stop_producer()
        except Exception:  # noqa: BLE001
            pass

    def reset(self):
        self._stop_producer()
        self._stop = threading.Event()
        order = list(self.records)
        if self.shuffle:
            random.shuffle(order)
        self._q = queue.Queue(maxsize=self.prefetch)
        self._thread = threading.Thread(target=self._producer, args=(order,), daemon=True)
        self._thread.start()

    def next(self):
        b = self._q.get()
        if b is None:
            self._q.put(None)
            raise StopIteration
        return b
