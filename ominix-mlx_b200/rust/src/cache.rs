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
