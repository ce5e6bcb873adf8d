"""Language-model text data sets (reference: python/mxnet/gluon/contrib/data/text.py ``WikiText2`` / ``WikiText103``).

A segment file (``wiki.train.tokens`` / ``wiki.valid.tokens`` / ``wiki.test.tokens``) is tokenised on whitespace, every line
gets an ``<eos>`` token, the token stream is mapped through a ``contrib.text.vocab.Vocabulary`` (built from the file unless
one is passed in) and cut into rows of ``seq_len`` tokens; item ``i`` is ``(data[i], label[i])`` with the label stream
shifted by one token.  The reference downloads the archive when the file is missing; here the files are read from
``root`` (or from the ``wikitext-*-v1.zip`` archive if that is what was staged there) and a missing file is an error that
names the expected path -- there is no outbound network on the training hosts this framework targets.
"""
import io
import os
import zipfile

import numpy as np

from ....contrib import text as _text
from ...data import dataset as _dataset
from .... import ndarray as nd

__all__ = ["WikiText2", "WikiText103"]

EOS_TOKEN = "<eos>"


class _WikiText(_dataset.Dataset):
    _namespace = None
    _archive = None
    _files = {"train": "wiki.train.tokens", "validation": "wiki.valid.tokens", "test": "wiki.test.tokens"}

    def __init__(self, root, segment="train", vocab=None, seq_len=35):
        if segment not in self._files:
            raise ValueError("segment must be one of %s" % sorted(self._files))
        self._root = os.path.expanduser(root)
        self._segment, self._vocab, self._seq_len = segment, vocab, int(seq_len)
        self._counter = None
        self._get_data()

    @property
    def vocabulary(self):
        return self._vocab

    @property
    def frequencies(self):
        return self._counter

    def _locate(self):
        name = self._files[self._segment]
        path = os.path.join(self._root, name)
        if os.path.exists(path):
            return path
        arc = os.path.join(self._root, self._archive)
        if os.path.exists(arc):                      # the published archive, staged by hand: unpack the token files
            with zipfile.ZipFile(arc, "r") as zf:
                for member in zf.namelist():
                    base = os.path.basename(member)
                    if base:
                        with zf.open(member) as src, open(os.path.join(self._root, base), "wb") as dst:
                            dst.write(src.read())
            if os.path.exists(path):
                return path
        raise RuntimeError("%s: %s not found (and no %s next to it); this host has no network access to download the "
                           "data set -- place the file there" % (type(self).__name__, path, self._archive))

    def _get_data(self):
        with io.open(self._locate(), "r", encoding="utf8") as f:
            content = f.read()
        lines = [ln.strip().split() for ln in content.splitlines()]
        tokens = [t for ln in lines if ln for t in ln + [EOS_TOKEN]]
        if self._counter is None:
            self._counter = _text.utils.count_tokens_from_str(content)
        if self._vocab is None:
            self._vocab = _text.vocab.Vocabulary(counter=self._counter, reserved_tokens=[EOS_TOKEN])
        ids = np.asarray(self._vocab.to_indices(tokens), dtype=np.int32)
        n = (len(ids) - 1) // self._seq_len * self._seq_len
        self._data = nd.array(ids[:n].reshape(-1, self._seq_len), dtype="int32")
        self._label = nd.array(ids[1:n + 1].reshape(-1, self._seq_len), dtype="int32")

    def __getitem__(self, idx):
        return self._data[idx], self._label[idx]

    def __len__(self):
        return self._label.shape[0]


class WikiText2(_WikiText):
    """WikiText-2 word-level language modelling data set (Merity et al. 2016), 2M training tokens."""
    _namespace, _archive = "wikitext-2", "wikitext-2-v1.zip"

    def __init__(self, root=os.path.join("~", ".mxnet", "datasets", "wikitext-2"), segment="train", vocab=None, seq_len=35):
        super(WikiText2, self).__init__(root, segment, vocab, seq_len)


class WikiText103(_WikiText):
    """WikiText-103: the 103M-token version."""
    _namespace, _archive = "wikitext-103", "wikitext-103-v1.zip"

    def __init__(self, root=os.path.join("~", ".mxnet", "datasets", "wikitext-103"), segment="train", vocab=None, seq_len=35):
        super(WikiText103, self).__init__(root, segment, vocab, seq_len)
