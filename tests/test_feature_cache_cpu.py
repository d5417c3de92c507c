"""FeatureCache bookkeeping with a stub extractor: misses are featurised once and in one batched
call, hits never touch the loader, LRU eviction, duplicates inside a batch, failure leaves no half-filled entries.
The cache contains no arithmetic, so a CPU stub exercises all of it; the GPU test runs it over the real extractor."""
import numpy as np
import pytest
import torch

from whisper_context_biasing_b200.feature_cache import FeatureCache, N_FRAMES


class StubExtractor:
    feature_size = 4
    device = torch.device("cpu")

    def __init__(self):
        self.calls = []

    def extract_host(self, clips):
        self.calls.append(len(clips))
        # "features" = the clip's first sample broadcast: enough to tell entries apart
        return torch.stack([torch.full((self.feature_size, N_FRAMES), float(c[0])) for c in clips])


def _loader(log):
    def load(key):
        log.append(key)
        return np.array([float(key)], dtype=np.float32)
    return load


def test_hits_misses_and_single_batched_call():
    ex, log = StubExtractor(), []
    item = 4 * N_FRAMES * 4
    cache = FeatureCache(ex, capacity_bytes=10 * item)
    a = cache.get_many([3, 5, 3, 7], _loader(log))
    assert a.shape == (4, 4, N_FRAMES) and a.dtype == torch.float32
    assert [float(a[i, 0, 0]) for i in range(4)] == [3.0, 5.0, 3.0, 7.0]
    assert log == [3, 5, 7] and ex.calls == [3]              # duplicates loaded once, one extractor call
    b = cache.get_many([7, 3], _loader(log))
    assert log == [3, 5, 7] and ex.calls == [3]              # pure hits: no loader, no extractor
    assert [float(b[i, 0, 0]) for i in range(2)] == [7.0, 3.0]
    st = cache.stats()
    assert st["entries"] == 3 and st["misses"] == 3 and st["hits"] == 2 and st["evictions"] == 0
    assert st["bytes"] == 3 * item


def test_lru_eviction_keeps_what_the_call_needs():
    ex, log = StubExtractor(), []
    item = 4 * N_FRAMES * 4
    cache = FeatureCache(ex, capacity_bytes=3 * item)
    cache.get_many([1, 2, 3], _loader(log))
    cache.get_many([1], _loader(log))                        # 1 becomes most recent: LRU order 2, 3, 1
    out = cache.get_many([4, 1], _loader(log))               # evicts 2, never 1 (needed by this call)
    assert [float(out[i, 0, 0]) for i in range(2)] == [4.0, 1.0]
    assert 2 not in cache and 1 in cache and 3 in cache and 4 in cache
    assert cache.stats()["evictions"] == 1
    with pytest.raises(RuntimeError):
        cache.get_many([10, 11, 12, 13], _loader(log))       # more distinct clips than slots
    assert len(cache) <= 3


def test_failure_leaves_no_entries_and_half_precision_storage():
    ex = StubExtractor()
    item16 = 4 * N_FRAMES * 2
    cache = FeatureCache(ex, capacity_bytes=4 * item16, dtype=torch.float16)
    assert cache.slots == 4

    def bad(key):
        raise IOError("decode failed")

    with pytest.raises(IOError):
        cache.get_many([1, 2], bad)
    assert len(cache) == 0 and len(cache._free) == 4
    out = cache.get_many([1, 2], _loader([]))
    assert out.dtype == torch.float32 and float(out[1, 0, 0]) == 2.0
    cache.clear()
    assert len(cache) == 0
    with pytest.raises(ValueError):
        FeatureCache(ex, capacity_bytes=10)
