#!/usr/bin/env python
"""Rebuild the .idx file of a RecordIO file (reference: tools/rec2idx.py): key = running record number."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mxb200 import recordio  # noqa: E402

if __name__ == "__main__":
    if len(sys.argv) != 3:
        sys.exit("usage: rec2idx.py file.rec file.idx")
    rd = recordio.MXRecordIO(sys.argv[1], "r")
    n = 0
    with open(sys.argv[2], "w") as f:
        while True:
            pos = rd.tell()
            if rd.read() is None:
                break
            f.write("%d\t%d\n" % (n, pos))
            n += 1
    rd.close()
    print("indexed %d records" % n)
