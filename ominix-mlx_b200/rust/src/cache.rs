//! KV caches with the reference's trait and growth contract (mlx-rs-core/src/cache.rs).
use std::sync::Arc;

use crate::array::{check, Array, Stream};
use crate::error::{Exception, Result};
use crate::ffi;

/// Trait for key-value caches used in attention (cache.rs:7-20) -- unchanged.
pub trait KeyValueCache {
    fn offset(&self) -> i32;
    fn max_size(&self) -> Option<i32>;
    fn update_and_fetch(&mut self, keys: Array, values: Array) -> std::result::Result<(Array, Array), Exception>;
    fn reset(&mut self) {}
}

impl<T: KeyValueCache> KeyValueCache for &'_ mut T {
    fn offset(&self) -> i32 {
        T::offset(self)
    }
    fn max_size(&self) -> Option<i32> {
        T::max_size(self)
    }
    fn update_and_fetch(&mut self, keys: Array, values: Array) -> std::result::Result<(Array, Array), Exception> {
        T::update_and_fetch(self, keys, values)
    }
    fn reset(&mut self) {
        T::reset(self)
    }
}

struct Handle(ffi::omx_kv_cache, bool /* concat */);
unsafe impl Send for Handle {}
unsafe impl Sync for Handle {}
impl Drop for Handle {
    fn drop(&mut self) {
        unsafe {
            if self.1 { ffi::omx_concat_kv_cache_free(self.0) } else { ffi::omx_kv_cache_free(self.0) };
        }
    }
}

fn empty_desc() -> ffi::omx_array {
    ffi::omx_array { data: std::ptr::null_mut(), dtype: 0, ndim: 0, shape: [0; 8], strides: [0; 8] }
}

/// Step-based KV cache with pre-allocation (cache.rs:92-195); device-resident, growth by the
/// reference rule, physical buffer doubling underneath.
#[derive(Clone)]
pub struct KVCache {
    h: Arc<Handle>,
    stream: Stream,
}

impl Default for KVCache {
    fn default() -> Self {
        Self::new()
    }
}

impl KVCache {
    pub fn new() -> Self {
        Self::with_step(256)
    }
    pub fn with_step(step: i32) -> Self {
        let mut h = ffi::omx_kv_cache { ctx: std::ptr::null_mut() };
        check(unsafe { ffi::omx_kv_cache_new(&mut h, step) }).expect("omx_kv_cache_new");
        Self { h: Arc::new(Handle(h, false)), stream: Stream::default() }
    }
    /// Stream the cache's copies run on (default: the default stream).
    pub fn on_stream(mut self, s: Stream) -> Self {
        self.stream = s;
        self
    }
    pub fn raw(&self) -> ffi::omx_kv_cache {
        self.h.0
    }
    /// CUDA-graph decode loop (include/omx_attn.h): pin the buffers and the cache-owned split-K scratch for
    /// positions [0, max_rows); addresses stay fixed until the cache outgrows them.
    pub fn prepare_graph(&mut self, max_rows: i32, n_q_heads: i32) -> Result<()> {
        check(unsafe { ffi::omx_kv_cache_prepare_graph(self.h.0, max_rows, n_q_heads, self.stream.0) })
    }
    /// Host bookkeeping for `n` rows appended by dynamic-position launches / graph replays.
    pub fn advance(&mut self, n: i32) -> Result<()> {
        check(unsafe { ffi::omx_kv_cache_advance(self.h.0, n, self.stream.0) })
    }
    pub(crate) fn keepalive(&self) -> Arc<dyn std::any::Any + Send + Sync> {
        self.h.clone()
    }
}

impl KeyValueCache for KVCache {
    fn offset(&self) -> i32 {
        let mut n = 0;
        unsafe { ffi::omx_kv_cache_offset(self.h.0, &mut n) };
        n
    }
    fn max_size(&self) -> Option<i32> {
        None
    }
    fn reset(&mut self) {
        unsafe { ffi::omx_kv_cache_reset(self.h.0) };
    }
    fn update_and_fetch(&mut self, keys: Array, values: Array) -> Result<(Array, Array)> {
        let (mut ko, mut vo) = (empty_desc(), empty_desc());
        check(unsafe {
            ffi::omx_kv_cache_update_and_fetch(self.h.0, keys.as_ptr(), values.as_ptr(), &mut ko, &mut vo, self.stream.0)
        })?;
        Ok((Array::from_desc(ko, self.h.clone()), Array::from_desc(vo, self.h.clone())))
    }
}

/// Simple concatenation-based KV cache (cache.rs:45-85).
#[derive(Clone)]
pub struct ConcatKeyValueCache {
    h: Arc<Handle>,
    stream: Stream,
}

impl Default for ConcatKeyValueCache {
    fn default() -> Self {
        Self::new()
    }
}

impl ConcatKeyValueCache {
    pub fn new() -> Self {
        let mut h = ffi::omx_kv_cache { ctx: std::ptr::null_mut() };
        check(unsafe { ffi::omx_concat_kv_cache_new(&mut h) }).expect("omx_concat_kv_cache_new");
        Self { h: Arc::new(Handle(h, true)), stream: Stream::default() }
    }
}

impl KeyValueCache for ConcatKeyValueCache {
    fn offset(&self) -> i32 {
        let mut n = 0;
        unsafe { ffi::omx_concat_kv_cache_offset(self.h.0, &mut n) };
        n
    }
    fn max_size(&self) -> Option<i32> {
        None
    }
    fn update_and_fetch(&mut self, keys: Array, values: Array) -> Result<(Array, Array)> {
        let (mut ko, mut vo) = (empty_desc(), empty_desc());
        check(unsafe {
            ffi::omx_concat_kv_cache_update_and_fetch(self.h.0, keys.as_ptr(), values.as_ptr(), &mut ko, &mut vo,
                                                      self.stream.0)
        })?;
        Ok((Array::from_desc(ko, self.h.clone()), Array::from_desc(vo, self.h.clone())))
    }
}

// ------------------------------------------------------------------------------------------------ paged
struct PagedHandle(ffi::omx_paged_kv_cache);
unsafe impl Send for PagedHandle {}
unsafe impl Sync for PagedHandle {}
impl Drop for PagedHandle {
    fn drop(&mut self) {
        unsafe { ffi::omx_paged_kv_cache_free(self.0) };
    }
}

/// The `KeyValueCache` contract over a page pool (include/omx_attn.h "paged KV cache"; north_star "paged KV
/// append"): pool `[n_pages][Hkv][64][D]` allocated once, block table, per-sequence lengths.  Growth takes a page
/// id from the free list -- no reallocation, no copy (the reference concatenates a fresh block every 256 tokens,
/// mlx-rs-core/src/cache.rs:141-181).  `update_and_fetch` appends to EVERY sequence and returns MATERIALISED
/// `[B, Hkv, offset, D]` arrays; `update` appends without materialising.
#[derive(Clone)]
pub struct PagedKVCache {
    h: Arc<PagedHandle>,
    batch: i32,
    stream: Stream,
}

impl PagedKVCache {
    pub fn new(batch: i32, n_kv_heads: i32, head_dim: i32, dtype: crate::array::Dtype, n_pages: i64,
               max_pages_per_seq: i32) -> Result<Self> {
        let mut h = ffi::omx_paged_kv_cache { ctx: std::ptr::null_mut() };
        check(unsafe {
            ffi::omx_paged_kv_cache_new(&mut h, batch, n_kv_heads, head_dim, head_dim, dtype as i32, n_pages,
                                        max_pages_per_seq)
        })?;
        Ok(Self { h: Arc::new(PagedHandle(h)), batch, stream: Stream::default() })
    }
    pub fn on_stream(mut self, s: Stream) -> Self {
        self.stream = s;
        self
    }
    pub fn raw(&self) -> ffi::omx_paged_kv_cache {
        self.h.0
    }
    /// Rows stored per sequence (host mirror); -1 = released slot.
    pub fn lengths(&self) -> Result<Vec<i32>> {
        let mut l = vec![0i32; self.batch as usize];
        check(unsafe { ffi::omx_paged_kv_cache_lengths(self.h.0, l.as_mut_ptr()) })?;
        Ok(l)
    }
    pub fn free_pages(&self) -> Result<i64> {
        let mut n = 0i64;
        check(unsafe { ffi::omx_paged_kv_cache_free_pages(self.h.0, &mut n) })?;
        Ok(n)
    }
    /// `KeyValueCache::reset` for one sequence: length 0, pages back to the free list.
    pub fn reset_slot(&mut self, slot: i32) -> Result<()> {
        check(unsafe { ffi::omx_paged_kv_cache_reset(self.h.0, slot, self.stream.0) })
    }
    /// Mark a slot inactive: the fused decode step skips it until `reset_slot`.
    pub fn release(&mut self, slot: i32) -> Result<()> {
        check(unsafe { ffi::omx_paged_kv_cache_release(self.h.0, slot, self.stream.0) })
    }
    /// Pre-assign pages for `rows_ahead` more rows per active sequence (no host-side allocation in the next
    /// `rows_ahead` fused steps: capturable into a CUDA graph, an even number of steps per capture).
    pub fn reserve(&mut self, rows_ahead: i32) -> Result<()> {
        check(unsafe { ffi::omx_paged_kv_cache_reserve(self.h.0, rows_ahead, self.stream.0) })
    }
    pub fn trim(&mut self, n: i32) -> Result<()> {
        check(unsafe { ffi::omx_paged_kv_cache_trim(self.h.0, n, self.stream.0) })
    }
    /// Host mirror <- device lengths (after graph replays of the fused step).
    pub fn sync_lengths(&mut self) -> Result<()> {
        check(unsafe { ffi::omx_paged_kv_cache_sync_lengths(self.h.0, self.stream.0) })
    }
    /// Append `[B, Hkv, n, D]` to every sequence without materialising the fetched views.
    pub fn update(&mut self, keys: &Array, values: &Array) -> Result<()> {
        check(unsafe {
            ffi::omx_paged_kv_cache_update_and_fetch(self.h.0, keys.as_ptr(), values.as_ptr(), std::ptr::null_mut(),
                                                     std::ptr::null_mut(), self.stream.0)
        })
    }
    /// Ragged prefill: `[1, Hkv, n, D]` rows appended to one sequence.
    pub fn append_slot(&mut self, slot: i32, keys: &Array, values: &Array) -> Result<()> {
        check(unsafe { ffi::omx_paged_kv_cache_append_slot(self.h.0, slot, keys.as_ptr(), values.as_ptr(), self.stream.0) })
    }
    /// Materialised `[B, Hkv, offset, D]` keys / values (rows past a shorter sequence's end read +0.0).
    pub fn fetch(&mut self) -> Result<(Array, Array)> {
        let (mut ko, mut vo) = (empty_desc(), empty_desc());
        check(unsafe { ffi::omx_paged_kv_cache_fetch(self.h.0, &mut ko, &mut vo, self.stream.0) })?;
        Ok((Array::from_desc(ko, self.h.clone()), Array::from_desc(vo, self.h.clone())))
    }
    pub(crate) fn keepalive(&self) -> Arc<dyn std::any::Any + Send + Sync> {
        self.h.clone()
    }
}

impl KeyValueCache for PagedKVCache {
    /// The longest sequence (== the common offset when the batch advances in lockstep, as in the reference).
    fn offset(&self) -> i32 {
        let mut n = 0;
        unsafe { ffi::omx_paged_kv_cache_offset(self.h.0, &mut n) };
        n
    }
    fn max_size(&self) -> Option<i32> {
        None
    }
    fn reset(&mut self) {
        unsafe { ffi::omx_paged_kv_cache_reset(self.h.0, -1, self.stream.0) };
    }
    fn update_and_fetch(&mut self, keys: Array, values: Array) -> Result<(Array, Array)> {
        let (mut ko, mut vo) = (empty_desc(), empty_desc());
        check(unsafe {
            ffi::omx_paged_kv_cache_update_and_fetch(self.h.0, keys.as_ptr(), values.as_ptr(), &mut ko, &mut vo,
                                                     self.stream.0)
        })?;
        Ok((Array::from_desc(ko, self.h.clone()), Array::from_desc(vo, self.h.clone())))
    }
}
