"""contrib.text.embedding (reference: python/mxnet/contrib/text/embedding.py): token embeddings indexed by a vocabulary.
Pre-trained GloVe / fastText files cannot be downloaded here (no network); ``CustomEmbedding`` loads any local
``token<delim>v1<delim>v2...`` file and ``CompositeEmbedding`` concatenates embeddings over one vocabulary."""
import io
import logging

import numpy as np

from ... import ndarray as nd
from . import vocab

_REG = {}


def register(embedding_cls):
    _REG[embedding_cls.__name__.lower()] = embedding_cls
    return embedding_cls


def create(embedding_name, **kwargs):
    if embedding_name.lower() not in _REG:
        raise KeyError("Cannot find `embedding_name` %s. Valid names: %s" % (embedding_name, ", ".join(sorted(_REG))))
    return _REG[embedding_name.lower()](**kwargs)


def get_pretrained_file_names(embedding_name=None):
    names = {"glove": ["glove.6B.50d.txt", "glove.6B.100d.txt", "glove.6B.200d.txt", "glove.6B.300d.txt", "glove.42B.300d.txt",
                       "glove.840B.300d.txt", "glove.twitter.27B.25d.txt"],
             "fasttext": ["wiki.en.vec", "wiki.simple.vec", "crawl-300d-2M.vec"]}
    return names if embedding_name is None else names[embedding_name.lower()]


class _TokenEmbedding(vocab.Vocabulary):
    pretrained_file_name_sha1 = {}

    @classmethod
    def _check_pretrained_file_names(cls, pretrained_file_name):
        """Raise KeyError if the file name is not one this embedding class knows how to fetch."""
        embedding_name = cls.__name__.lower()
        if pretrained_file_name not in cls.pretrained_file_name_sha1:
            raise KeyError("Cannot find pretrained file %s for token embedding %s. Valid pretrained files for embedding %s: %s"
                           % (pretrained_file_name, embedding_name, embedding_name, ", ".join(cls.pretrained_file_name_sha1.keys())))

    def __init__(self, **kwargs):
        super(_TokenEmbedding, self).__init__(**kwargs)
        self._vec_len = 0
        self._idx_to_vec = None

    def _load_embedding(self, path, elem_delim, init_unknown_vec, encoding="utf8"):
        vecs, toks = [], []
        seen = set()
        loaded_unknown_vec = None
        with io.open(path, "r", encoding=encoding) as f:
            for n, line in enumerate(f):
                parts = line.rstrip().split(elem_delim)
                assert len(parts) > 1, "At line %d of the pre-trained text embedding file: the data format of the " \
                    "pre-trained token embedding file %s is unexpected." % (n + 1, path)
                if len(parts) == 2:                # a 1-dimensional "vector": the fastText header "count dim", skipped
                    logging.warning("At line %d of %s: token %s with a 1-dimensional vector is likely a header and is skipped",
                                    n + 1, path, parts[0])
                    continue
                tok, vals = parts[0], parts[1:]
                if tok == self.unknown_token:
                    # the file's own vector for the unknown token wins over init_unknown_vec (embedding.py:232-260)
                    if loaded_unknown_vec is None:
                        loaded_unknown_vec = np.asarray(vals, dtype=np.float32)
                    continue
                if tok in seen:
                    logging.warning("duplicate token %r in %s skipped (first occurrence kept)", tok, path)
                    continue
                seen.add(tok)
                assert not self._vec_len or len(vals) == self._vec_len, \
                    "At line %d of the pre-trained token embedding file: the dimension of token %s is %d but the dimension " \
                    "of previous tokens is %d. Dimensions of all the tokens must be the same." % (n + 1, tok, len(vals), self._vec_len)
                self._vec_len = self._vec_len or len(vals)
                toks.append(tok)
                vecs.append(np.asarray(vals, dtype=np.float32))
        for t in toks:
            if t not in self._token_to_idx:
                self._idx_to_token.append(t)
                self._token_to_idx[t] = len(self._idx_to_token) - 1
        mat = np.zeros((len(self._idx_to_token), self._vec_len), dtype=np.float32)
        if loaded_unknown_vec is not None and len(loaded_unknown_vec) == self._vec_len:
            mat[0] = loaded_unknown_vec
        else:
            mat[0] = init_unknown_vec(shape=self._vec_len).asnumpy() if init_unknown_vec is not None else 0
        for t, v in zip(toks, vecs):
            mat[self._token_to_idx[t]] = v
        self._idx_to_vec = nd.array(mat)

    def _build_for_vocabulary(self, vocabulary, embeddings):
        # look the vocabulary's tokens up in every source embedding FIRST (a source may be this very object, whose own
        # token table is replaced by the vocabulary's below)
        # row 0 -- the vocabulary's unknown token, whatever it is called -- takes each source's UNKNOWN vector
        cols = []
        for emb in embeddings:
            m = emb.get_vecs_by_tokens(list(vocabulary.idx_to_token)).asnumpy().copy()
            m[0] = emb.idx_to_vec[0].asnumpy()
            cols.append(m)
        self._idx_to_token = list(vocabulary.idx_to_token)
        self._token_to_idx = dict(vocabulary.token_to_idx)
        self._unknown_token, self._reserved_tokens = vocabulary.unknown_token, vocabulary.reserved_tokens
        mat = np.concatenate(cols, axis=1)
        self._vec_len = mat.shape[1]
        self._idx_to_vec = nd.array(mat)

    @property
    def vec_len(self):
        return self._vec_len

    @property
    def idx_to_vec(self):
        return self._idx_to_vec

    def get_vecs_by_tokens(self, tokens, lower_case_backup=False):
        single = not isinstance(tokens, list)
        toks = [tokens] if single else tokens
        if lower_case_backup:
            idx = [self._token_to_idx.get(t, self._token_to_idx.get(t.lower(), 0)) for t in toks]
        else:
            idx = [self._token_to_idx.get(t, 0) for t in toks]
        vecs = nd.Embedding(nd.array(idx), self._idx_to_vec, input_dim=self._idx_to_vec.shape[0], output_dim=self._vec_len)
        return vecs[0] if single else vecs

    def update_token_vectors(self, tokens, new_vectors):
        assert self._idx_to_vec is not None, "The property `idx_to_vec` has not been properly set."
        toks = [tokens] if not isinstance(tokens, list) else tokens
        assert isinstance(new_vectors, nd.NDArray) and len(new_vectors.shape) in (1, 2), \
            "`new_vectors` must be a 1-D or 2-D NDArray."
        nv = new_vectors.asnumpy()
        nv = nv.reshape(1, -1) if nv.ndim == 1 else nv
        assert nv.shape == (len(toks), self._vec_len), \
            "The length of new_vectors must be equal to the number of tokens and the width of new_vectors must be equal " \
            "to the dimension of embeddings of the glossary."
        mat = self._idx_to_vec.asnumpy()
        for t, v in zip(toks, nv):
            if t not in self._token_to_idx:
                raise ValueError("Token %s is unknown. To update the embedding vector for an unknown token, please specify "
                                 "it explicitly as the `unknown_token` %s" % (t, self.unknown_token))
            mat[self._token_to_idx[t]] = v
        self._idx_to_vec = nd.array(mat)


@register
class CustomEmbedding(_TokenEmbedding):
    """Embedding loaded from a local text file ``token<elem_delim>v1<elem_delim>v2...`` (embedding.py:659-718)."""

    def __init__(self, pretrained_file_path, elem_delim=" ", encoding="utf8", init_unknown_vec=nd.zeros, vocabulary=None, **kwargs):
        super(CustomEmbedding, self).__init__(**kwargs)
        self._load_embedding(pretrained_file_path, elem_delim, init_unknown_vec, encoding)
        if vocabulary is not None:
            self._build_for_vocabulary(vocabulary, [self])


class _Pretrained(CustomEmbedding):
    def __init__(self, pretrained_file_name=None, embedding_root=None, **kwargs):
        import os
        root = os.path.expanduser(embedding_root or os.path.join("~", ".mxnet", "embeddings"))
        path = os.path.join(root, self.__class__.__name__.lower(), pretrained_file_name or self._default)
        if not os.path.exists(path):
            raise IOError("pre-trained file %s not found and cannot be downloaded in this environment (no network); place "
                          "it there or use CustomEmbedding" % path)
        super(_Pretrained, self).__init__(path, elem_delim=" ", **kwargs)


@register
class GloVe(_Pretrained):
    _default = "glove.840B.300d.txt"


@register
class FastText(_Pretrained):
    _default = "wiki.simple.vec"


class CompositeEmbedding(_TokenEmbedding):
    """Concatenation of several token embeddings over one vocabulary (embedding.py:720-770)."""

    def __init__(self, vocabulary, token_embeddings):
        if not isinstance(token_embeddings, list):
            token_embeddings = [token_embeddings]
        super(CompositeEmbedding, self).__init__()
        self._build_for_vocabulary(vocabulary, token_embeddings)
