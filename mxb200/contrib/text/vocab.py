"""contrib.text.vocab.Vocabulary (reference: python/mxnet/contrib/text/vocab.py): index 0 is the unknown token, then
the reserved tokens, then counter keys by descending frequency (ties broken alphabetically)."""
import collections


class Vocabulary(object):
    def __init__(self, counter=None, most_freq_count=None, min_freq=1, unknown_token="<unk>", reserved_tokens=None):
        assert min_freq > 0, "`min_freq` must be set to a positive value."
        if reserved_tokens is not None:
            rs = set(reserved_tokens)
            assert unknown_token not in rs, "`reserved_token` cannot contain `unknown_token`."
            assert len(rs) == len(reserved_tokens), "`reserved_tokens` cannot contain duplicate reserved tokens."
        self._unknown_token = unknown_token
        self._reserved_tokens = list(reserved_tokens) if reserved_tokens is not None else None
        self._idx_to_token = [unknown_token] + (self._reserved_tokens or [])
        self._token_to_idx = {t: i for i, t in enumerate(self._idx_to_token)}
        if counter is not None:
            assert isinstance(counter, collections.Counter), "`counter` must be an instance of collections.Counter."
            special = set(self._idx_to_token)
            pairs = sorted(counter.items(), key=lambda x: x[0])
            pairs.sort(key=lambda x: x[1], reverse=True)
            cap = len(special) + (len(counter) if most_freq_count is None else most_freq_count)
            for tok, freq in pairs:
                if freq < min_freq or len(self._idx_to_token) == cap:
                    break
                if tok not in special:
                    self._idx_to_token.append(tok)
                    self._token_to_idx[tok] = len(self._idx_to_token) - 1

    def __len__(self):
        return len(self._idx_to_token)

    @property
    def token_to_idx(self):
        return self._token_to_idx

    @property
    def idx_to_token(self):
        return self._idx_to_token

    @property
    def unknown_token(self):
        return self._unknown_token

    @property
    def reserved_tokens(self):
        return self._reserved_tokens

    def to_indices(self, tokens):
        single = not isinstance(tokens, list)
        toks = [tokens] if single else tokens
        idx = [self._token_to_idx.get(t, 0) for t in toks]
        return idx[0] if single else idx

    def to_tokens(self, indices):
        single = not isinstance(indices, list)
        idx = [indices] if single else indices
        for i in idx:
            if not isinstance(i, int) or i < 0 or i >= len(self._idx_to_token):
                raise ValueError("Token index %s in the provided `indices` is invalid." % str(i))
        toks = [self._idx_to_token[i] for i in idx]
        return toks[0] if single else toks
