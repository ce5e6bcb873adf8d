"""Bucketed sentence iterator (reference: python/mxnet/rnn/io.py).

Provenance: adapted from Apache MXNet (incubating) 1.3.1 ``python/mxnet/rnn/io.py`` (Apache License 2.0).  The public API, argument
names and control flow follow the original so that user scripts keep working; the tensor maths runs on torch /
the mxb200 kernels.  This file is API-compatibility glue, not original B200 work ("Provenance").
"""
import bisect
import random

import numpy as np

from ..io import DataBatch, DataDesc, DataIter
from .. import ndarray

__all__ = ["encode_sentences", "BucketSentenceIter"]


def encode_sentences(sentences, vocab=None, invalid_label=-1, invalid_key="\n", start_label=0, unknown_token=None):
    """token lists -> integer id lists; a new vocabulary is built when ``vocab`` is None (io.py:30-81)."""
    idx = start_label
    if vocab is None:
        vocab = {invalid_key: invalid_label}
        new_vocab = True
    else:
        new_vocab = False
    res = []
    for sent in sentences:
        coded = []
        for word in sent:
            if word not in vocab:
                assert new_vocab or unknown_token, "Unknown token %s" % word
                if idx == invalid_label:
                    idx += 1
                if unknown_token:
                    word = unknown_token
                if word not in vocab:
                    vocab[word] = idx
                    idx += 1
            coded.append(vocab[word])
        res.append(coded)
    return res, vocab


class BucketSentenceIter(DataIter):
    """Pads every sentence to the smallest bucket that fits it and yields batches bucket by bucket (``bucket_key`` =
    sequence length), for BucketingModule; the label is the input shifted by one token (io.py:84-205)."""

    def __init__(self, sentences, batch_size, buckets=None, invalid_label=-1, data_name="data", label_name="softmax_label",
                 dtype="float32", layout="NT"):
        super(BucketSentenceIter, self).__init__()
        if not buckets:
            buckets = [i for i, j in enumerate(np.bincount([len(s) for s in sentences])) if j >= batch_size]
        buckets.sort()
        ndiscard = 0
        self.data = [[] for _ in buckets]
        for sent in sentences:
            buck = bisect.bisect_left(buckets, len(sent))
            if buck == len(buckets):
                ndiscard += 1
                continue
            buff = np.full((buckets[buck],), invalid_label, dtype=dtype)
            buff[:len(sent)] = sent
            self.data[buck].append(buff)
        self.data = [np.asarray(i, dtype=dtype) for i in self.data]
        self.batch_size, self.buckets = batch_size, buckets
        self.data_name, self.label_name, self.dtype = data_name, label_name, dtype
        self.invalid_label, self.nddata, self.ndlabel = invalid_label, [], []
        self.major_axis = layout.find("N")
        self.layout = layout
        self.default_bucket_key = max(buckets)
        shape = (batch_size, self.default_bucket_key) if self.major_axis == 0 else (self.default_bucket_key, batch_size)
        self.provide_data = [DataDesc(self.data_name, shape, layout=layout)]
        self.provide_label = [DataDesc(self.label_name, shape, layout=layout)]
        self.idx = []
        for i, buck in enumerate(self.data):
            self.idx.extend([(i, j) for j in range(0, len(buck) - batch_size + 1, batch_size)])
        self.curr_idx = 0
        self.reset()

    def reset(self):
        self.curr_idx = 0
        random.shuffle(self.idx)
        for buck in self.data:
            np.random.shuffle(buck)
        self.nddata, self.ndlabel = [], []
        for buck in self.data:
            label = np.empty_like(buck)
            if buck.size:
                label[:, :-1] = buck[:, 1:]
                label[:, -1] = self.invalid_label
            self.nddata.append(ndarray.array(buck, dtype=self.dtype))
            self.ndlabel.append(ndarray.array(label, dtype=self.dtype))

    def next(self):
        if self.curr_idx == len(self.idx):
            raise StopIteration
        i, j = self.idx[self.curr_idx]
        self.curr_idx += 1
        if self.major_axis == 1:
            data = self.nddata[i][j:j + self.batch_size].T
            label = self.ndlabel[i][j:j + self.batch_size].T
        else:
            data = self.nddata[i][j:j + self.batch_size]
            label = self.ndlabel[i][j:j + self.batch_size]
        return DataBatch([data], [label], pad=0, bucket_key=self.buckets[i],
                         provide_data=[DataDesc(self.data_name, data.shape, layout=self.layout)],
                         provide_label=[DataDesc(self.label_name, label.shape, layout=self.layout)])
