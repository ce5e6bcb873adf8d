"""kvstore helpers (reference: src/kvstore/kvstore_utils.{cc,cu}: UniqueImpl = sort + unique of row ids)."""
import torch


def unique_rows(ids):
    """Sorted unique int64 row ids (GPU: radix sort + unique inside torch.unique / CUB)."""
    return torch.unique(ids.reshape(-1).to(torch.int64), sorted=True)
