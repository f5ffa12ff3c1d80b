"""Mirror of mlx-rs-core/src/cache.rs: KeyValueCache, KVCache (step 256), ConcatKeyValueCache.
The growth / offset logic lives in the native library (csrc/kv_cache.cu); these classes own a
handle and turn the fetched descriptors into zero-copy tensor views."""
import ctypes

from . import _lib
from .array import desc, ref, stream_ptr, view


class KeyValueCache:
    """trait KeyValueCache (cache.rs:7-20)."""

    def offset(self):
        raise NotImplementedError

    def max_size(self):
        raise NotImplementedError

    def update_and_fetch(self, keys, values):
        raise NotImplementedError

    def reset(self):  # default: does nothing (cache.rs:17-19)
        pass


class KVCache(KeyValueCache):
    """cache.rs:92-195.  KVCache() == with_step(256)."""

    _new = "omx_kv_cache_new"
    _free = "omx_kv_cache_free"
    _offset = "omx_kv_cache_offset"
    _update = "omx_kv_cache_update_and_fetch"

    def __init__(self, step=256):
        self._h = _lib.OmxKVCache()
        self._make(step)

    def _make(self, step):
        _lib.check(_lib.lib().omx_kv_cache_new(ctypes.byref(self._h), int(step)))

    @classmethod
    def with_step(cls, step):
        return cls(step)

    def __del__(self):
        try:
            if getattr(self, "_h", None) is not None and self._h.ctx:
                getattr(_lib.lib(), self._free)(self._h)
                self._h.ctx = None
        except Exception:
            pass

    def offset(self):
        n = ctypes.c_int(0)
        _lib.check(getattr(_lib.lib(), self._offset)(self._h, ctypes.byref(n)))
        return n.value

    def max_size(self):
        return None

    def reset(self):
        _lib.check(_lib.lib().omx_kv_cache_reset(self._h))

    def reserve(self, rows):
        """Extension (not in the reference): pre-size the allocation; logical growth is unchanged."""
        _lib.check(_lib.lib().omx_kv_cache_reserve(self._h, int(rows)))

    def trim(self, n):
        """Extension (mlx-lm KVCache.trim): drop the last min(n, offset) rows, return how many."""
        t = ctypes.c_int(0)
        _lib.check(_lib.lib().omx_kv_cache_trim(self._h, int(n), ctypes.byref(t)))
        return t.value

    def prepare_graph(self, max_rows, n_q_heads, stream=None):
        """Extension (CUDA-graph decode loop, include/omx_attn.h): pin the buffers and the cache-owned
        split-K scratch for positions [0, max_rows); addresses stay fixed until the cache outgrows them."""
        _lib.check(_lib.lib().omx_kv_cache_prepare_graph(self._h, int(max_rows), int(n_q_heads), stream_ptr(stream)))

    def advance(self, n=1, stream=None):
        """Host bookkeeping for n rows appended by dynamic-position launches / graph replays."""
        _lib.check(_lib.lib().omx_kv_cache_advance(self._h, int(n), stream_ptr(stream)))

    def update_and_fetch(self, keys, values, stream=None):
        k, v = desc(keys), desc(values)
        ko, vo = _lib.OmxArray(), _lib.OmxArray()
        _lib.check(getattr(_lib.lib(), self._update)(self._h, ref(k), ref(v), ref(ko), ref(vo), stream_ptr(stream)))
        return view(ko, self, keys.device), view(vo, self, values.device)

    def state(self):
        """Whole backing buffers [B, Hkv, cap, D] (== self.keys / self.values in the reference)."""
        import torch
        ko, vo = _lib.OmxArray(), _lib.OmxArray()
        _lib.check(_lib.lib().omx_kv_cache_state(self._h, ref(ko), ref(vo)))
        dev = torch.device("cuda", torch.cuda.current_device())
        return view(ko, self, dev), view(vo, self, dev)

    @property
    def handle(self):
        return self._h


class ConcatKeyValueCache(KVCache):
    """cache.rs:45-85."""

    _new = "omx_concat_kv_cache_new"
    _free = "omx_concat_kv_cache_free"
    _offset = "omx_concat_kv_cache_offset"
    _update = "omx_concat_kv_cache_update_and_fetch"

    def __init__(self):
        self._h = _lib.OmxKVCache()
        _lib.check(_lib.lib().omx_concat_kv_cache_new(ctypes.byref(self._h)))

    def reset(self):  # trait default: no-op
        pass
