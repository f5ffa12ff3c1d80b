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


class PagedKVCache(KeyValueCache):
    """The KeyValueCache contract over a page pool (csrc/paged_kv.cu; north_star "paged KV append"):
    pool [n_pages][Hkv][64][D] allocated once, block table [batch][max_pages_per_seq], per-sequence lengths.
    Growth takes a page id from the free list -- no reallocation, no copy (the reference concatenates a
    fresh block every 256 tokens, cache.rs:141-181).  update_and_fetch appends to EVERY sequence and, like the
    reference, returns [B,Hkv,offset,D] tensors -- here materialised copies, produced only when `fetch=True`."""

    def __init__(self, batch, n_kv_heads, head_dim, dtype, n_pages, max_pages_per_seq=None, head_dim_v=None):
        from .array import _DT
        self._h = _lib.OmxPagedKVCache()
        self.batch, self.n_kv_heads, self.head_dim, self.dtype = int(batch), int(n_kv_heads), int(head_dim), dtype
        self.n_pages = int(n_pages)
        self.max_pages_per_seq = int(max_pages_per_seq or n_pages)
        _lib.check(_lib.lib().omx_paged_kv_cache_new(
            ctypes.byref(self._h), self.batch, self.n_kv_heads, self.head_dim, int(head_dim_v or head_dim), _DT[dtype],
            self.n_pages, self.max_pages_per_seq))

    PAGE_ROWS = 64

    def __del__(self):
        try:
            if getattr(self, "_h", None) is not None and self._h.ctx:
                _lib.lib().omx_paged_kv_cache_free(self._h)
                self._h.ctx = None
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def offset(self):
        """Longest sequence (== the common offset when the batch advances in lockstep, as in the reference)."""
        n = ctypes.c_int(0)
        _lib.check(_lib.lib().omx_paged_kv_cache_offset(self._h, ctypes.byref(n)))
        return n.value

    def max_size(self):
        return None

    def lengths(self):
        """Rows stored per sequence (host mirror); -1 = released slot."""
        out = (ctypes.c_int32 * self.batch)()
        _lib.check(_lib.lib().omx_paged_kv_cache_lengths(self._h, out))
        return list(out)

    def free_pages(self):
        n = ctypes.c_int64(0)
        _lib.check(_lib.lib().omx_paged_kv_cache_free_pages(self._h, ctypes.byref(n)))
        return n.value

    def reset(self, slot=-1, stream=None):
        """KeyValueCache::reset (cache.rs:130-132) for one sequence or all: length 0, pages returned."""
        _lib.check(_lib.lib().omx_paged_kv_cache_reset(self._h, int(slot), stream_ptr(stream)))

    def release(self, slot, stream=None):
        """Mark a slot inactive: the fused decode step skips it until reset(slot)."""
        _lib.check(_lib.lib().omx_paged_kv_cache_release(self._h, int(slot), stream_ptr(stream)))

    def reserve(self, rows_ahead, stream=None):
        _lib.check(_lib.lib().omx_paged_kv_cache_reserve(self._h, int(rows_ahead), stream_ptr(stream)))

    def trim(self, n, stream=None):
        _lib.check(_lib.lib().omx_paged_kv_cache_trim(self._h, int(n), stream_ptr(stream)))

    def sync_lengths(self, stream=None):
        """Host mirror <- device lengths (after capturing / replaying fused steps in a CUDA graph)."""
        _lib.check(_lib.lib().omx_paged_kv_cache_sync_lengths(self._h, stream_ptr(stream)))

    def update_and_fetch(self, keys, values, stream=None, fetch=True):
        k, v = desc(keys), desc(values)
        ko, vo = _lib.OmxArray(), _lib.OmxArray()
        _lib.check(_lib.lib().omx_paged_kv_cache_update_and_fetch(
            self._h, ref(k), ref(v), ref(ko) if fetch else None, ref(vo) if fetch else None, stream_ptr(stream)))
        if not fetch:
            return None
        return view(ko, self, keys.device), view(vo, self, values.device)

    def append_slot(self, slot, keys, values, stream=None):
        """Ragged prefill: [1,Hkv,n,D] rows appended to one sequence."""
        k, v = desc(keys), desc(values)
        _lib.check(_lib.lib().omx_paged_kv_cache_append_slot(self._h, int(slot), ref(k), ref(v), stream_ptr(stream)))

    def fetch(self, stream=None):
        """Materialised [B,Hkv,offset,D] keys / values (rows past a shorter sequence's end are +0.0)."""
        import torch
        ko, vo = _lib.OmxArray(), _lib.OmxArray()
        _lib.check(_lib.lib().omx_paged_kv_cache_fetch(self._h, ref(ko), ref(vo), stream_ptr(stream)))
        dev = torch.device("cuda", torch.cuda.current_device())
        return view(ko, self, dev), view(vo, self, dev)

    def pages(self):
        """(k_pool_ptr, v_pool_ptr, block_table as a list of per-sequence page lists) -- introspection."""
        kp, vp = ctypes.c_void_p(), ctypes.c_void_p()
        bt = ctypes.POINTER(ctypes.c_int32)()
        mp = ctypes.c_int(0)
        _lib.check(_lib.lib().omx_paged_kv_cache_pages(self._h, ctypes.byref(kp), ctypes.byref(vp), ctypes.byref(bt),
                                                       ctypes.byref(mp)))
        lens = self.lengths()
        table = [[bt[b * mp.value + t] for t in range((max(lens[b], 0) + 63) // 64)] for b in range(self.batch)]
        return kp.value, vp.value, table
