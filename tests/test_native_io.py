"""Native RecordIO reader / batch loader (csrc/host/recordio.cc) and the GPU image pipeline's host logic
(mxb200/io/gpu_pipeline.py).  Oracle: the pure-python reader in mxb200/recordio.py and OpenCV.
Reference tests of the same surface: tests/python/unittest/test_recordio.py, test_io.py::test_ImageRecordIter*."""
import os
import struct

import numpy as np
import pytest

import mxb200 as mx
from mxb200 import recordio
from mxb200.io.gpu_pipeline import GPUImageRecordIter, RecordFile, crop_resize_reference, WINDOW

cv2 = pytest.importorskip("cv2")


def _smooth_image(rng, h, w):
    # low-frequency content: JPEG round trips within a few levels
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    img = np.stack([127 + 100 * np.sin(xx / rng.uniform(8, 30) + rng.uniform(0, 3)) * np.cos(yy / rng.uniform(8, 30)),
                    127 + 100 * np.sin((xx + yy) / rng.uniform(10, 40)),
                    127 + 100 * np.cos(yy / rng.uniform(5, 25) + rng.uniform(0, 3))], axis=2)
    return np.clip(img, 0, 255).astype(np.uint8)


def _make_rec(tmp_path, n=23, seed=0, label_width=1, fmt=".jpg", sizes=None):
    rng = np.random.default_rng(seed)
    rec, idx = str(tmp_path / "data.rec"), str(tmp_path / "data.idx")
    w = recordio.MXIndexedRecordIO(idx, rec, "w")
    imgs, labels = [], []
    for i in range(n):
        h, wd = sizes[i] if sizes else (int(rng.integers(40, 90)), int(rng.integers(40, 90)))
        img = _smooth_image(rng, h, wd)
        lab = float(i % 7) if label_width == 1 else [float(i), float(i) / 2, 3.0][:label_width]
        header = recordio.IRHeader(0, lab, i, 0)
        w.write_idx(i, recordio.pack_img(header, img[:, :, ::-1], quality=95, img_fmt=fmt))   # pack_img takes BGR like cv2
        imgs.append(img)
        labels.append(lab)
    w.close()
    return rec, idx, imgs, labels


def test_native_reader_matches_python_reader_including_split_records(tmp_path):
    rec = str(tmp_path / "raw.rec")
    w = recordio.MXRecordIO(rec, "w")
    magic = struct.pack("<I", 0xced7230a)
    payloads = [b"hello", b"", b"x" * 1001, b"ab" + magic + b"cd", magic, b"12" + magic + magic + b"3", magic + b"tail" + magic,
                bytes(range(256)) * 5]
    for p in payloads:
        w.write(p)
    w.close()
    f = RecordFile(rec)
    assert len(f) == len(payloads)
    for i, p in enumerate(payloads):
        assert f.read(i) == p, i
        assert f.info(i)[0] == len(p)
    # the python reader is the oracle for the framing
    r = recordio.MXRecordIO(rec, "r")
    for p in payloads:
        assert r.read() == p
    r.close()
    f.close()


def test_native_reader_uses_idx_file_and_rejects_garbage(tmp_path):
    rec, idx, _imgs, _l = _make_rec(tmp_path, n=9)
    f = RecordFile(rec, idx)
    g = RecordFile(rec)
    assert len(f) == len(g) == 9
    r = recordio.MXIndexedRecordIO(idx, rec, "r")
    for i in range(9):
        assert f.read(i) == g.read(i) == r.read_idx(i)
        assert f.info(i)[1] == i
    bad = tmp_path / "bad.rec"
    bad.write_bytes(b"\x00" * 64)
    with pytest.raises(mx.MXNetError):
        RecordFile(str(bad))
    with pytest.raises(mx.MXNetError):
        RecordFile(str(tmp_path / "missing.rec"))


def _collect(it):
    out = []
    for b in it:
        out.append((b.data[0].asnumpy(), b.label[0].asnumpy(), b.pad, np.asarray(b.index)))
    return out


def test_loader_visits_every_record_once_with_padding_and_sharding(tmp_path):
    rec, idx, imgs, labels = _make_rec(tmp_path, n=23, sizes=[(48, 48)] * 23)
    it = GPUImageRecordIter(rec, (3, 48, 48), 5, path_imgidx=idx, threads=3, prefetch=2, layout="NHWC")
    batches = _collect(it)
    assert len(batches) == 5 and it.num_batches() == 5
    ids = np.concatenate([b[3] for b in batches])
    assert [b[2] for b in batches] == [0, 0, 0, 0, 2]
    assert list(ids[:23]) == list(range(23)) and list(ids[23:]) == [0, 1]          # round_batch wraps to the start
    lab = np.concatenate([b[1] for b in batches])
    assert np.allclose(lab[:23], labels)
    # decoded pixels: identity window (48x48 source, 48x48 output) equals the OpenCV decode of the same bytes
    x = np.concatenate([b[0] for b in batches])
    for i in range(23):
        ref = cv2.imdecode(np.frombuffer(recordio.unpack(RecordFile(rec, idx).read(i))[1], np.uint8), cv2.IMREAD_COLOR)[:, :, ::-1]
        assert np.abs(x[i] - ref.astype(np.float32)).max() < 1e-3
        assert np.abs(x[i] - imgs[i].astype(np.float32)).mean() < 4.0              # and is the image we wrote, up to JPEG loss
    # second epoch after reset() yields the same thing; shards partition the file
    it.reset()
    again = _collect(it)
    assert all(np.array_equal(a[0], b[0]) for a, b in zip(batches, again))
    seen = []
    for part in range(3):
        sh = GPUImageRecordIter(rec, (3, 48, 48), 4, path_imgidx=idx, num_parts=3, part_index=part, round_batch=False, layout="NHWC")
        got = _collect(sh)
        seen += [int(v) for b in got for v in b[3]]
        assert all(b[2] == 0 for b in got)
        sh.close()
    assert sorted(seen) == list(range(23))
    it.close()


def test_loader_shuffle_is_seeded_and_repartition_follows_membership(tmp_path):
    rec, idx, _imgs, _labels = _make_rec(tmp_path, n=20, sizes=[(40, 40)] * 20)
    a = GPUImageRecordIter(rec, (3, 40, 40), 4, path_imgidx=idx, shuffle=True, seed=5)
    b = GPUImageRecordIter(rec, (3, 40, 40), 4, path_imgidx=idx, shuffle=True, seed=5)
    ia = np.concatenate([x[3] for x in _collect(a)])
    ib = np.concatenate([x[3] for x in _collect(b)])
    assert np.array_equal(ia, ib) and sorted(ia) == list(range(20)) and list(ia) != list(range(20))
    a.reset()
    ia2 = np.concatenate([x[3] for x in _collect(a)])
    assert sorted(ia2) == list(range(20)) and list(ia2) != list(ia)                # a new permutation each epoch
    # elastic: 1 worker -> 2 workers at the epoch boundary
    a.set_partition(2, 1)
    a.reset()
    ia3 = np.concatenate([x[3] for x in _collect(a)])
    assert sorted(set(int(v) for v in ia3)) == list(range(10, 20)) and len(ia3) == 12   # 3 batches of 4, the last one wraps
    a.close()
    b.close()


def test_resize_center_crop_matches_opencv_resize_then_crop(tmp_path):
    rec, idx, _imgs, _labels = _make_rec(tmp_path, n=6, seed=3, sizes=[(60, 90), (90, 60), (64, 64), (50, 120), (77, 53), (100, 100)])
    it = GPUImageRecordIter(rec, (3, 32, 32), 6, path_imgidx=idx, resize=40, mean=(10, 20, 30), std=(2, 4, 8))
    x = _collect(it)[0][0]                                                          # NCHW float32
    f = RecordFile(rec, idx)
    for i in range(6):
        src = cv2.imdecode(np.frombuffer(recordio.unpack(f.read(i))[1], np.uint8), cv2.IMREAD_COLOR)[:, :, ::-1]
        h, w = src.shape[:2]
        nh, nw = (40 * h // w, 40) if h > w else (40, 40 * w // h)
        r = cv2.resize(src, (nw, nh), interpolation=cv2.INTER_LINEAR).astype(np.float32)
        y0, x0 = (nh - 32) // 2, (nw - 32) // 2
        ref = (r[y0:y0 + 32, x0:x0 + 32] - np.array([10, 20, 30], np.float32)) / np.array([2, 4, 8], np.float32)
        got = x[i].transpose(1, 2, 0)
        # OpenCV interpolates with 11-bit fixed-point weights and rounds to uint8: one grey level of slack (pre-normalisation)
        assert np.abs((got - ref) * np.array([2, 4, 8], np.float32)).max() <= 1.01, i
    it.close()


def test_random_windows_stay_inside_the_image_and_mirror_flips(tmp_path):
    rec, idx, _imgs, _labels = _make_rec(tmp_path, n=16, seed=1)
    it = GPUImageRecordIter(rec, (3, 24, 24), 16, path_imgidx=idx, random_resized_crop=True, min_random_area=0.08, max_random_area=1.0,
                            min_aspect_ratio=0.75, max_aspect_ratio=1.333, rand_mirror=True, seed=7)
    w = np.array([it.rec and 0] * 16)
    ws = np.random.default_rng(0).integers(40, 90, 16).astype(np.int32)
    hs = np.random.default_rng(1).integers(40, 90, 16).astype(np.int32)
    win = it._windows(16, ws, hs)
    assert win.dtype == WINDOW
    assert (win["x0"] >= 0).all() and (win["y0"] >= 0).all()
    assert (win["x0"] + win["cw"] <= ws + 1e-3).all() and (win["y0"] + win["ch"] <= hs + 1e-3).all()
    area = win["cw"] * win["ch"] / (ws * hs)
    assert (area > 0.07).all() and (area <= 1.0 + 1e-5).all()
    assert set(np.unique(win["mirror"])) <= {0, 1} and 0 < win["mirror"].sum() < 16
    del w
    # mirror = horizontal flip of the un-mirrored output
    img = _smooth_image(np.random.default_rng(2), 50, 70)
    one = np.zeros(1, WINDOW)
    one["w"], one["h"], one["x0"], one["y0"], one["cw"], one["ch"] = 70, 50, 5.5, 3.25, 40.0, 30.0
    a = crop_resize_reference(img, one[0], 24, 24)
    one["mirror"] = 1
    b = crop_resize_reference(img, one[0], 24, 24)
    assert np.allclose(a[:, ::-1], b)
    it.close()


def test_png_payloads_and_multi_labels(tmp_path):
    rec, idx, imgs, _labels = _make_rec(tmp_path, n=5, label_width=3, fmt=".png", sizes=[(33, 47)] * 5)
    it = GPUImageRecordIter(rec, (3, 33, 47), 5, path_imgidx=idx, label_width=3, layout="NHWC")
    (x, y, pad, _ids), = _collect(it)
    assert pad == 0 and y.shape == (5, 3)
    assert np.allclose(y[:, 0], np.arange(5)) and np.allclose(y[:, 1], np.arange(5) / 2) and np.allclose(y[:, 2], 3.0)
    for i in range(5):
        assert np.array_equal(x[i].astype(np.uint8), imgs[i])                       # PNG is lossless
    it.close()
