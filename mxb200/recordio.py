"""RecordIO reader/writer (reference: python/mxnet/recordio.py, dmlc-core recordio format).

Record layout: u32 magic 0xced7230a | u32 (cflag << 29 | length) | payload | pad to 4 bytes.  cflag: 0 whole
record, 1 first / 2 middle / 3 last part of a record that contained the magic word.  IRHeader packs
(flag, label, id, id2) in front of image bytes; flag > 0 means ``flag`` float labels follow the header.

Provenance: adapted from Apache MXNet (incubating) 1.3.1 ``python/mxnet/recordio.py`` (Apache License 2.0).  The public API, argument
names and control flow follow the original so that user scripts keep working; the tensor maths runs on torch /
the mxb200 kernels.  This file is API-compatibility glue, not original B200 work ("Provenance").
"""
import ctypes
import numbers
import os
import struct
from collections import namedtuple

import numpy as np

__all__ = ["MXRecordIO", "MXIndexedRecordIO", "IRHeader", "pack", "unpack", "pack_img", "unpack_img"]

_MAGIC = 0xced7230a
_MAGIC_BYTES = struct.pack("<I", _MAGIC)


class MXRecordIO(object):
    def __init__(self, uri, flag):
        self.uri = uri
        self.flag = flag
        self.pid = None
        self.is_open = False
        self.open()

    def open(self):
        if self.flag == "w":
            self.handle = open(self.uri, "wb")
            self.writable = True
        elif self.flag == "r":
            self.handle = open(self.uri, "rb")
            self.writable = False
        else:
            raise ValueError("Invalid flag %s" % self.flag)
        self.pid = os.getpid()
        self.is_open = True

    def __del__(self):
        self.close()

    def __getstate__(self):
        is_open = self.is_open
        self.close()
        d = dict(self.__dict__)
        d["is_open"] = is_open
        d.pop("handle", None)
        return d

    def __setstate__(self, d):
        self.__dict__ = d
        is_open = d["is_open"]
        self.is_open = False
        self.handle = None
        if is_open:
            self.open()

    def _check_pid(self, allow_reset=False):
        if not self.pid == os.getpid():
            if allow_reset:
                self.reset()
            else:
                raise RuntimeError("Forbidden operation in multiple processes")

    def close(self):
        if not self.is_open:
            return
        self.handle.close()
        self.is_open = False
        self.pid = None

    def tell(self):
        """Byte offset of the next record (MXRecordIOWriterTell / ReaderTell in the reference C API)."""
        return self.handle.tell()

    def reset(self):
        self.close()
        self.open()

    def write(self, buf):
        assert self.writable
        self._check_pid(allow_reset=False)
        # split at embedded magic words (dmlc RecordIOWriter::WriteRecord)
        parts = []
        start = 0
        while True:
            pos = buf.find(_MAGIC_BYTES, start)
            while pos != -1 and pos % 4 != 0:
                pos = buf.find(_MAGIC_BYTES, pos + 1)
            if pos == -1:
                parts.append(buf[start:])
                break
            parts.append(buf[start:pos])
            start = pos + 4
        n = len(parts)
        for i, p in enumerate(parts):
            cflag = 0 if n == 1 else (1 if i == 0 else (3 if i == n - 1 else 2))
            self.handle.write(struct.pack("<II", _MAGIC, (cflag << 29) | len(p)))
            self.handle.write(p)
            pad = (4 - len(p) % 4) % 4
            if pad:
                self.handle.write(b"\x00" * pad)

    def read(self):
        assert not self.writable
        self._check_pid(allow_reset=True)
        out = b""
        while True:
            head = self.handle.read(8)
            if len(head) < 8:
                return out if out else None
            magic, lrec = struct.unpack("<II", head)
            if magic != _MAGIC:
                raise IOError("Invalid RecordIO file: bad magic")
            cflag, length = lrec >> 29, lrec & ((1 << 29) - 1)
            data = self.handle.read(length)
            pad = (4 - length % 4) % 4
            if pad:
                self.handle.read(pad)
            if cflag == 0:
                return data
            out += data
            if cflag == 3:
                return out
            out += _MAGIC_BYTES


class MXIndexedRecordIO(MXRecordIO):
    def __init__(self, idx_path, uri, flag, key_type=int):
        self.idx_path = idx_path
        self.idx = {}
        self.keys = []
        self.key_type = key_type
        self.fidx = None
        super(MXIndexedRecordIO, self).__init__(uri, flag)

    def open(self):
        super(MXIndexedRecordIO, self).open()
        self.idx = {}
        self.keys = []
        self.fidx = open(self.idx_path, self.flag)
        if not self.writable:
            for line in iter(self.fidx.readline, ""):
                line = line.strip().split("\t")
                key = self.key_type(line[0])
                self.idx[key] = int(line[1])
                self.keys.append(key)

    def close(self):
        if not self.is_open:
            return
        super(MXIndexedRecordIO, self).close()
        self.fidx.close()

    def __getstate__(self):
        d = super(MXIndexedRecordIO, self).__getstate__()
        d["fidx"] = None
        return d

    def seek(self, idx):
        assert not self.writable
        self._check_pid(allow_reset=True)
        self.handle.seek(self.idx[idx])

    def tell(self):
        assert self.writable
        return self.handle.tell()

    def read_idx(self, idx):
        self.seek(idx)
        return self.read()

    def write_idx(self, idx, buf):
        key = self.key_type(idx)
        pos = self.tell()
        self.write(buf)
        self.fidx.write("%s\t%d\n" % (str(key), pos))
        self.idx[key] = pos
        self.keys.append(key)


IRHeader = namedtuple("HEADER", ["flag", "label", "id", "id2"])
_IR_FORMAT = "IfQQ"
_IR_SIZE = struct.calcsize(_IR_FORMAT)


def pack(header, s):
    header = IRHeader(*header)
    if isinstance(header.label, numbers.Number):
        header = header._replace(flag=0)
    else:
        label = np.asarray(header.label, dtype=np.float32)
        header = header._replace(flag=label.size, label=0)
        s = label.tobytes() + s
    return struct.pack(_IR_FORMAT, *header) + s


def unpack(s):
    header = IRHeader(*struct.unpack(_IR_FORMAT, s[:_IR_SIZE]))
    s = s[_IR_SIZE:]
    if header.flag > 0:
        header = header._replace(label=np.frombuffer(s, np.float32, header.flag))
        s = s[header.flag * 4:]
    return header, s


def unpack_img(s, iscolor=-1):
    header, s = unpack(s)
    import cv2
    img = cv2.imdecode(np.frombuffer(s, dtype=np.uint8), iscolor)
    return header, img


def pack_img(header, img, quality=95, img_fmt=".jpg"):
    import cv2
    jpg_formats = [".JPG", ".JPEG"]
    png_formats = [".PNG"]
    encode_params = None
    if img_fmt.upper() in jpg_formats:
        encode_params = [cv2.IMWRITE_JPEG_QUALITY, quality]
    elif img_fmt.upper() in png_formats:
        encode_params = [cv2.IMWRITE_PNG_COMPRESSION, quality]
    ret, buf = cv2.imencode(img_fmt, img, encode_params)
    assert ret, "failed to encode image"
    return pack(header, buf.tobytes())
