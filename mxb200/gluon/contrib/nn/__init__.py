"""gluon.contrib.nn (reference: python/mxnet/gluon/contrib/nn/basic_layers.py)."""
from ...block import Block, HybridBlock
from ...nn.basic_layers import BatchNorm, HybridSequential, Sequential

__all__ = ["Concurrent", "HybridConcurrent", "Identity", "SparseEmbedding", "SyncBatchNorm"]


class Concurrent(Sequential):
    """Feeds the same input to every child and concatenates their outputs along ``axis`` (basic_layers.py:29-59)."""

    def __init__(self, axis=-1, prefix=None, params=None):
        super(Concurrent, self).__init__(prefix=prefix, params=params)
        self.axis = axis

    def forward(self, x):
        from .... import ndarray as nd
        return nd.concat(*[blk(x) for blk in self._children.values()], dim=self.axis)


class HybridConcurrent(HybridSequential):
    def __init__(self, axis=-1, prefix=None, params=None):
        super(HybridConcurrent, self).__init__(prefix=prefix, params=params)
        self.axis = axis

    def hybrid_forward(self, F, x):
        return F.concat(*[blk(x) for blk in self._children.values()], dim=self.axis)


class Identity(HybridBlock):
    """Pass-through block, handy as a branch of a (Hybrid)Concurrent (basic_layers.py:95-114)."""

    def hybrid_forward(self, F, x):
        return x


class SparseEmbedding(Block):
    """Embedding whose gradient is row-sparse in the reference (basic_layers.py:116-160).  The fused optimizers here
    are dense, so the weight keeps the default storage; the layer exists for API compatibility."""

    def __init__(self, input_dim, output_dim, dtype="float32", weight_initializer=None, **kwargs):
        super(SparseEmbedding, self).__init__(**kwargs)
        self._kwargs = {"input_dim": input_dim, "output_dim": output_dim, "dtype": dtype}
        self.weight = self.params.get("weight", shape=(input_dim, output_dim), init=weight_initializer, dtype=dtype)

    def forward(self, x):
        from .... import ndarray as nd
        return nd.Embedding(x, self.weight._raw_data(x.context), **self._kwargs)

    def __repr__(self):
        return "{name}({input_dim} -> {output_dim}, {dtype})".format(name=self.__class__.__name__, **self._kwargs)


class SyncBatchNorm(BatchNorm):
    """BatchNorm with statistics reduced over all data-parallel devices (basic_layers.py:163-240).  ``num_devices`` is
    kept for signature parity; here the reduction spans the torch.distributed process group (one process per GPU) and
    happens on the device (ndarray/op_extra.py::_contrib_SyncBatchNorm)."""

    def __init__(self, in_channels=0, num_devices=None, momentum=0.9, epsilon=1e-5, center=True, scale=True,
                 use_global_stats=False, beta_initializer="zeros", gamma_initializer="ones",
                 running_mean_initializer="zeros", running_variance_initializer="ones", **kwargs):
        super(SyncBatchNorm, self).__init__(1, momentum, epsilon, center, scale, use_global_stats, beta_initializer,
                                            gamma_initializer, running_mean_initializer, running_variance_initializer,
                                            in_channels, **kwargs)
        self._sync_kwargs = {"eps": epsilon, "momentum": momentum, "fix_gamma": not scale, "use_global_stats": use_global_stats,
                             "ndev": num_devices or 1, "key": self.prefix}

    def hybrid_forward(self, F, x, gamma, beta, running_mean, running_var):
        return F.contrib.SyncBatchNorm(x, gamma, beta, running_mean, running_var, name="fwd", **self._sync_kwargs)
