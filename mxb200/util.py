"""General utilities (reference: python/mxnet/util.py)."""
import os


def makedirs(d):
    os.makedirs(d, exist_ok=True)
