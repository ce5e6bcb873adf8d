"""Device image pipeline on a real GPU: nvJPEG batched decode + fused crop / resize / mirror / normalise kernel
(csrc/cuda/jpeg.cu) behind mx.io.ImageRecordIter(ctx=mx.gpu(0)).  Oracles: OpenCV decode of the same bytes and the
numpy formula of the kernel (mxb200/io/gpu_pipeline.py::crop_resize_reference)."""
import numpy as np
import pytest
import torch

import mxb200 as mx
from mxb200 import recordio
from mxb200.io.gpu_pipeline import GPUImageRecordIter, RecordFile, crop_resize_reference

from test_native_io import _make_rec, _collect

pytestmark = pytest.mark.gpu
cv2 = pytest.importorskip("cv2")


def _cv_decode(rec, idx, i):
    f = RecordFile(rec, idx)
    a = cv2.imdecode(np.frombuffer(recordio.unpack(f.read(i))[1], np.uint8), cv2.IMREAD_COLOR)[:, :, ::-1]
    f.close()
    return a


@pytest.mark.parametrize("backend", [0, 2])
def test_nvjpeg_identity_window_matches_opencv_decode(tmp_path, backend):
    sizes = [(64, 64), (48, 80), (80, 48), (33, 47), (96, 96), (64, 64), (50, 70), (72, 40)]
    rec, idx, imgs, labels = _make_rec(tmp_path, n=8, sizes=sizes)
    for i, (h, w) in enumerate(sizes):
        it = GPUImageRecordIter(rec, (3, h, w), 8, path_imgidx=idx, layout="NHWC", ctx=mx.gpu(0), backend=backend)
        assert it.backend == backend
        (x, y, pad, ids), = _collect(it)
        assert x.shape == (8, h, w, 3) and pad == 0 and list(ids) == list(range(8))
        assert np.allclose(y, labels)
        ref = _cv_decode(rec, idx, i).astype(np.float32)
        # two JPEG decoders (different IDCT / chroma up-sampling): close, not identical
        d = np.abs(x[i] - ref)
        assert d.mean() < 3.5 and d.max() <= 32, (i, d.mean(), d.max())
        it.close()


def test_gpu_crop_resize_kernel_matches_numpy_formula(tmp_path):
    sizes = [(60, 90), (90, 60), (64, 64), (50, 120), (77, 53), (100, 100), (41, 43), (88, 66)]
    rec, idx, _imgs, _labels = _make_rec(tmp_path, n=8, seed=4, sizes=sizes)
    # decoded pixels of every image through an identity window (per-image size)
    decoded = []
    for i, (h, w) in enumerate(sizes):
        it = GPUImageRecordIter(rec, (3, h, w), 8, path_imgidx=idx, layout="NHWC", ctx=mx.gpu(0))
        decoded.append(np.rint(_collect(it)[0][0][i]).astype(np.uint8))
        it.close()
    mean, std = (123.68, 116.28, 103.53), (58.4, 57.1, 57.4)
    for kw in [dict(resize=48), dict(resize=48, rand_crop=True, rand_mirror=True),
               dict(random_resized_crop=True, min_random_area=0.08, max_random_area=1.0, min_aspect_ratio=0.75, max_aspect_ratio=1.333,
                    rand_mirror=True)]:
        for layout in ("NCHW", "NHWC"):
            it = GPUImageRecordIter(rec, (3, 32, 40), 8, path_imgidx=idx, layout=layout, ctx=mx.gpu(0), mean=mean, std=std, seed=11, **kw)
            # the windows the iterator drew for this batch: same generator state as a fresh iterator on the CPU
            cpu_it = GPUImageRecordIter(rec, (3, 32, 40), 8, path_imgidx=idx, layout=layout, mean=mean, std=std, seed=11, **kw)
            x = _collect(it)[0][0]
            cpu_it._rng = np.random.default_rng([11, 0, 0])
            win = cpu_it._windows(8, np.array([s[1] for s in sizes], np.int32), np.array([s[0] for s in sizes], np.int32))
            for i in range(8):
                ref = (crop_resize_reference(decoded[i], win[i], 32, 40) - np.array(mean, np.float32)) / np.array(std, np.float32)
                got = x[i] if layout == "NHWC" else x[i].transpose(1, 2, 0)
                assert np.abs(got - ref).max() < 2e-3, (kw, layout, i, np.abs(got - ref).max())
            it.close()
            cpu_it.close()


def test_gpu_iterator_bf16_nhwc_many_epochs_and_png_fallback(tmp_path):
    rec, idx, imgs, _labels = _make_rec(tmp_path, n=37, seed=2)
    it = mx.io.ImageRecordIter(path_imgrec=rec, path_imgidx=idx, data_shape=(3, 32, 32), batch_size=8, shuffle=True, rand_crop=True,
                               rand_mirror=True, resize=36, mean_r=123.68, mean_g=116.28, mean_b=103.53, std_r=58.4, std_g=57.1,
                               std_b=57.4, dtype="bfloat16", layout="NHWC", ctx=mx.gpu(0), preprocess_threads=3, prefetch_buffer=3)
    assert isinstance(it, GPUImageRecordIter)
    for epoch in range(4):
        seen, nb = [], 0
        for b in it:
            x = b.data[0]._t
            assert x.dtype == torch.bfloat16 and x.is_cuda and tuple(x.shape) == (8, 32, 32, 3)
            assert torch.isfinite(x.float()).all() and x.float().abs().max() < 4.0
            seen += [int(v) for v in b.index[:8 - b.pad]]
            nb += 1
        assert nb == 5 and sorted(seen) == list(range(37))
        it.reset()
    it.close()
    # PNG payloads take the host-decode branch and still land in the same arena / kernel
    (tmp_path / "png").mkdir()
    rec2, idx2, imgs2, _l = _make_rec(tmp_path / "png", n=4, fmt=".png", sizes=[(40, 56)] * 4)
    it = GPUImageRecordIter(rec2, (3, 40, 56), 4, path_imgidx=idx2, layout="NHWC", ctx=mx.gpu(0))
    x = _collect(it)[0][0]
    for i in range(4):
        assert np.array_equal(x[i].astype(np.uint8), imgs2[i])
    it.close()
