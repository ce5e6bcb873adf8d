#!/usr/bin/env python
"""Create image lists and RecordIO files (reference: tools/im2rec.py / im2rec.cc).

    python tools/im2rec.py --list prefix image_root          # write prefix.lst  (index \\t label \\t relative path)
    python tools/im2rec.py prefix image_root [--resize 256] [--quality 95] [--pass-through]   # prefix.lst -> prefix.rec/.idx
"""
import argparse
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mxb200 import recordio  # noqa: E402

EXTS = (".jpg", ".jpeg", ".png", ".bmp")


def list_images(root, recursive=True):
    cat, i = {}, 0
    for path, dirs, files in sorted(os.walk(root, followlinks=True)):
        dirs.sort()
        for f in sorted(files):
            if f.lower().endswith(EXTS):
                label = cat.setdefault(os.path.relpath(path, root), len(cat))
                yield i, os.path.relpath(os.path.join(path, f), root), label
                i += 1
        if not recursive:
            break


def write_list(prefix, root, shuffle=True, train_ratio=1.0):
    items = list(list_images(root))
    if shuffle:
        random.seed(100)
        random.shuffle(items)
    n_train = int(len(items) * train_ratio)
    chunks = [("", items)] if train_ratio >= 1.0 else [("_train", items[:n_train]), ("_val", items[n_train:])]
    for suffix, it in chunks:
        with open(prefix + suffix + ".lst", "w") as f:
            for idx, path, label in it:
                f.write("%d\t%f\t%s\n" % (idx, float(label), path))


def read_list(path):
    with open(path) as f:
        for line in f:
            p = line.strip().split("\t")
            if len(p) >= 3:
                yield int(p[0]), [float(v) for v in p[1:-1]], p[-1]


def make_rec(prefix, root, resize=0, quality=95, pass_through=False, color=1, encoding=".jpg"):
    rec = recordio.MXIndexedRecordIO(prefix + ".idx", prefix + ".rec", "w")
    n = 0
    for idx, labels, rel in read_list(prefix + ".lst"):
        full = os.path.join(root, rel)
        header = recordio.IRHeader(0, labels[0] if len(labels) == 1 else labels, idx, 0)
        if pass_through:
            with open(full, "rb") as f:
                rec.write_idx(idx, recordio.pack(header, f.read()))
        else:
            import cv2
            img = cv2.imread(full, color)
            if img is None:
                print("imread failed:", full)
                continue
            if resize:
                h, w = img.shape[:2]
                s = float(resize) / min(h, w)
                img = cv2.resize(img, (max(1, int(round(w * s))), max(1, int(round(h * s)))))
            rec.write_idx(idx, recordio.pack_img(header, img, quality=quality, img_fmt=encoding))
        n += 1
        if n % 1000 == 0:
            print("processed", n)
    rec.close()
    print("wrote %d records to %s.rec" % (n, prefix))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("prefix")
    ap.add_argument("root")
    ap.add_argument("--list", action="store_true")
    ap.add_argument("--no-shuffle", action="store_true")
    ap.add_argument("--train-ratio", type=float, default=1.0)
    ap.add_argument("--resize", type=int, default=0)
    ap.add_argument("--quality", type=int, default=95)
    ap.add_argument("--pass-through", action="store_true")
    ap.add_argument("--color", type=int, default=1)
    ap.add_argument("--encoding", default=".jpg")
    a = ap.parse_args()
    if a.list:
        write_list(a.prefix, a.root, not a.no_shuffle, a.train_ratio)
    else:
        make_rec(a.prefix, a.root, a.resize, a.quality, a.pass_through, a.color, a.encoding)
