"""Local store of pretrained model-zoo weights (reference: python/mxnet/gluon/model_zoo/model_store.py ``get_model_file`` /
``purge``).

The reference resolves ``<name>-<first 8 hex digits of the sha1>.params`` under ``root`` (default
``$MXNET_HOME/models``, i.e. ``~/.mxnet/models``) and downloads the zip from the Apache repository when the file is
missing or its checksum is wrong.  Here the lookup is the same -- any ``<name>-<hash8>.params`` (or ``<name>.params``)
in ``root`` is accepted, and when the file name carries a hash prefix the file's sha1 must start with it -- but a missing
file is an error that says where to put it: training clusters of this framework run without outbound network, so weights
are staged into the store by the deployment (``MXNET_GLUON_REPO`` may point ``gluon.utils.download`` at a mirror that is
reachable as a local path).
"""
import glob
import hashlib
import os

from ...base import MXNetError

__all__ = ["get_model_file", "purge", "short_hash"]


def _root(root=None):
    root = root or os.path.join(os.environ.get("MXNET_HOME", os.path.join("~", ".mxnet")), "models")
    return os.path.expanduser(root)


def _sha1(path):
    h = hashlib.sha1()
    with open(path, "rb") as f:
        for chunk in iter(lambda: f.read(1 << 20), b""):
            h.update(chunk)
    return h.hexdigest()


def short_hash(path):
    """First 8 hex digits of the file's sha1 (the suffix the store's file names carry)."""
    return _sha1(path)[:8]


def get_model_file(name, root=None):
    """Path of the parameter file of ``name`` in the local store; raises when it is not there (no download)."""
    root = _root(root)
    cands = sorted(glob.glob(os.path.join(root, name + "-*.params"))) + \
        [p for p in [os.path.join(root, name + ".params")] if os.path.exists(p)]
    bad = []
    for path in cands:
        stem = os.path.basename(path)[:-len(".params")]
        tag = stem[len(name) + 1:] if stem != name else ""
        if tag and len(tag) == 8 and all(c in "0123456789abcdef" for c in tag) and not _sha1(path).startswith(tag):
            bad.append(path)                       # truncated / corrupted copy: the name promises another checksum
            continue
        return path
    repo = os.environ.get("MXNET_GLUON_REPO", "")
    if repo and os.path.isdir(repo):               # a mirror mounted as a directory
        for path in sorted(glob.glob(os.path.join(repo, "gluon", "models", name + "-*.params")) +
                           glob.glob(os.path.join(repo, name + "-*.params"))):
            return path
    msg = "Pretrained model file for %r not found under %s (no network access to download it): copy %s-<hash>.params there" \
        % (name, root, name)
    if bad:
        msg += "; files with a checksum mismatch were ignored: %s" % ", ".join(bad)
    raise MXNetError(msg)


def purge(root=None):
    """Delete every parameter file of the local store."""
    root = _root(root)
    for f in glob.glob(os.path.join(root, "*.params")):
        os.remove(f)
