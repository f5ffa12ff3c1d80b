//! Raw declarations of include/omx_attn.h (what bindgen produces for mlx-c in mlx-sys).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

pub const OMX_MAX_NDIM: usize = 8;

/// Values equal mlx_dtype (mlx-c/mlx/c/array.h:37-52).
#[repr(i32)]
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum omx_dtype {
    Bool = 0,
    Int32 = 7,
    Float16 = 9,
    Float32 = 10,
    Bfloat16 = 12,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct omx_array {
    pub data: *mut c_void,
    pub dtype: i32,
    pub ndim: i32,
    pub shape: [i64; OMX_MAX_NDIM],
    pub strides: [i64; OMX_MAX_NDIM],
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct omx_optional_float {
    pub value: f32,
    pub has_value: bool,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct omx_kv_cache {
    pub ctx: *mut c_void,
}

pub const OMX_MAX_PEERS: usize = 8;

/// Handle of the paged KV cache (include/omx_attn.h: omx_paged_kv_cache).
#[repr(C)]
#[derive(Clone, Copy)]
pub struct omx_paged_kv_cache {
    pub ctx: *mut c_void,
}

/// Peer mappings of the head-sharded decode step (include/omx_attn.h: omx_peer_group).
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct omx_peer_group {
    pub world: i32,
    pub rank: i32,
    pub out: [*mut c_void; OMX_MAX_PEERS],
    pub flags: [*mut u32; OMX_MAX_PEERS],
}

/// Staging buffers of the data + flag exchange (include/omx_attn.h: omx_ll_group).
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct omx_ll_group {
    pub world: i32,
    pub rank: i32,
    pub staging: [*mut c_void; OMX_MAX_PEERS],
    pub seq: *mut u32,
}

pub type omx_stream = *mut c_void; // cudaStream_t
pub type omx_error_handler_func = Option<unsafe extern "C" fn(msg: *const c_char, data: *mut c_void)>;

extern "C" {
    pub fn omx_set_error_handler(handler: omx_error_handler_func, data: *mut c_void,
                                 dtor: Option<unsafe extern "C" fn(*mut c_void)>);
    pub fn omx_last_error() -> *const c_char;
    pub fn omx_version() -> c_int;
    pub fn omx_device_check(sm: *mut c_int) -> c_int;

    pub fn omx_fast_rms_norm(out: *const omx_array, x: *const omx_array, weight: *const omx_array, eps: f32,
                             s: omx_stream) -> c_int;
    pub fn omx_fast_rope(out: *const omx_array, x: *const omx_array, dims: c_int, traditional: bool,
                         base: omx_optional_float, scale: f32, offset: c_int, freqs: *const omx_array,
                         s: omx_stream) -> c_int;
    pub fn omx_fast_rope_dynamic(out: *const omx_array, x: *const omx_array, dims: c_int, traditional: bool,
                                 base: omx_optional_float, scale: f32, offset: *const omx_array,
                                 max_position: c_int, freqs: *const omx_array, s: omx_stream) -> c_int;
    pub fn omx_fast_scaled_dot_product_attention(out: *const omx_array, queries: *const omx_array,
                                                 keys: *const omx_array, values: *const omx_array,
                                                 scale: f32, mask_mode: *const c_char,
                                                 mask_arr: *const omx_array, sinks: *const omx_array,
                                                 s: omx_stream) -> c_int;

    pub fn omx_kv_cache_new(res: *mut omx_kv_cache, step: c_int) -> c_int;
    pub fn omx_kv_cache_free(c: omx_kv_cache) -> c_int;
    pub fn omx_kv_cache_offset(c: omx_kv_cache, offset: *mut c_int) -> c_int;
    pub fn omx_kv_cache_reset(c: omx_kv_cache) -> c_int;
    pub fn omx_kv_cache_update_and_fetch(c: omx_kv_cache, keys: *const omx_array, values: *const omx_array,
                                         keys_out: *mut omx_array, values_out: *mut omx_array,
                                         s: omx_stream) -> c_int;
    pub fn omx_kv_cache_state(c: omx_kv_cache, keys_buf: *mut omx_array, values_buf: *mut omx_array) -> c_int;
    pub fn omx_kv_cache_trim(c: omx_kv_cache, n: c_int, trimmed: *mut c_int) -> c_int;
    pub fn omx_kv_cache_reserve(c: omx_kv_cache, rows: c_int) -> c_int;
    pub fn omx_concat_kv_cache_new(res: *mut omx_kv_cache) -> c_int;
    pub fn omx_concat_kv_cache_free(c: omx_kv_cache) -> c_int;
    pub fn omx_concat_kv_cache_offset(c: omx_kv_cache, offset: *mut c_int) -> c_int;
    pub fn omx_concat_kv_cache_update_and_fetch(c: omx_kv_cache, keys: *const omx_array,
                                                values: *const omx_array, keys_out: *mut omx_array,
                                                values_out: *mut omx_array, s: omx_stream) -> c_int;

    pub fn omx_attn_decode_fused(out: *const omx_array, q: *const omx_array, k_new: *const omx_array,
                                 v_new: *const omx_array, cache: omx_kv_cache, rope_dims: c_int,
                                 traditional: bool, base: omx_optional_float, rope_scale: f32,
                                 freqs: *const omx_array, sm_scale: f32, keys_out: *mut omx_array,
                                 values_out: *mut omx_array, s: omx_stream) -> c_int;
    pub fn omx_attn_decode_fused_norm(out: *const omx_array, q: *const omx_array, k_new: *const omx_array,
                                      v_new: *const omx_array, cache: omx_kv_cache,
                                      q_norm_weight: *const omx_array, k_norm_weight: *const omx_array,
                                      norm_eps: f32, rope_dims: c_int, traditional: bool,
                                      base: omx_optional_float, rope_scale: f32, freqs: *const omx_array,
                                      sm_scale: f32, keys_out: *mut omx_array, values_out: *mut omx_array,
                                      s: omx_stream) -> c_int;
    // CUDA-graph decode loop (position read on the device; see include/omx_attn.h)
    pub fn omx_kv_cache_prepare_graph(c: omx_kv_cache, max_rows: c_int, n_q_heads: c_int, s: omx_stream) -> c_int;
    pub fn omx_kv_cache_advance(c: omx_kv_cache, n: c_int, s: omx_stream) -> c_int;
    pub fn omx_attn_decode_fused_dynamic(out: *const omx_array, q: *const omx_array, k_new: *const omx_array,
                                         v_new: *const omx_array, cache: omx_kv_cache,
                                         q_norm_weight: *const omx_array, k_norm_weight: *const omx_array,
                                         norm_eps: f32, rope_dims: c_int, traditional: bool,
                                         base: omx_optional_float, rope_scale: f32, sm_scale: f32,
                                         position: *const i32, s: omx_stream) -> c_int;
    pub fn omx_device_counter_add(counter: *mut i32, delta: c_int, s: omx_stream) -> c_int;
    pub fn omx_attn_prefill_fused(out: *const omx_array, q: *const omx_array, k_new: *const omx_array,
                                  v_new: *const omx_array, cache: omx_kv_cache, q_norm_weight: *const omx_array,
                                  k_norm_weight: *const omx_array, norm_eps: f32, rope_dims: c_int,
                                  traditional: bool, base: omx_optional_float, rope_scale: f32,
                                  freqs: *const omx_array, sm_scale: f32, mask_mode: *const c_char,
                                  mask_arr: *const omx_array, keys_out: *mut omx_array,
                                  values_out: *mut omx_array, s: omx_stream) -> c_int;
    pub fn omx_attn_decode_fused_sharded(out_full: *const omx_array, q: *const omx_array,
                                         k_new: *const omx_array, v_new: *const omx_array,
                                         cache: omx_kv_cache, rope_dims: c_int, traditional: bool,
                                         base: omx_optional_float, rope_scale: f32, freqs: *const omx_array,
                                         sm_scale: f32, peers: *const omx_peer_group, head_offset: c_int,
                                         s: omx_stream) -> c_int;
    pub fn omx_attn_decode_seqshard(partial: *const omx_array, q: *const omx_array, k_new: *const omx_array,
                                    v_new: *const omx_array, cache: omx_kv_cache, rope_dims: c_int,
                                    traditional: bool, base: omx_optional_float, rope_scale: f32,
                                    position: c_int, append: bool, sm_scale: f32,
                                    peers: *const omx_peer_group, s: omx_stream) -> c_int;
    pub fn omx_seqshard_merge(out: *const omx_array, partial: *const omx_array, peers: *const omx_peer_group,
                              expected: u32, s: omx_stream) -> c_int;
    pub fn omx_peer_wait(peers: *const omx_peer_group, expected: u32, s: omx_stream) -> c_int;
    pub fn omx_attn_decode_fused_sharded_sync(out_full: *const omx_array, q: *const omx_array,
                                              k_new: *const omx_array, v_new: *const omx_array,
                                              cache: omx_kv_cache, rope_dims: c_int, traditional: bool,
                                              base: omx_optional_float, rope_scale: f32,
                                              freqs: *const omx_array, sm_scale: f32,
                                              peers: *const omx_peer_group, head_offset: c_int,
                                              s: omx_stream) -> c_int;

    pub fn omx_attn_decode_fused_sharded_ll(out_full: *const omx_array, q: *const omx_array,
                                            k_new: *const omx_array, v_new: *const omx_array,
                                            cache: omx_kv_cache, rope_dims: c_int, traditional: bool,
                                            base: omx_optional_float, rope_scale: f32,
                                            freqs: *const omx_array, sm_scale: f32,
                                            group: *const omx_ll_group, head_offset: c_int,
                                            s: omx_stream) -> c_int;
    pub fn omx_ll_staging_bytes(world: c_int, b: i64, hq_local: i64, d: i64, dtype: c_int) -> usize;

    pub fn omx_paged_kv_cache_new(res: *mut omx_paged_kv_cache, batch: c_int, n_kv_heads: c_int, head_dim_k: c_int,
                                  head_dim_v: c_int, dtype: c_int, n_pages: i64, max_pages_per_seq: c_int) -> c_int;
    pub fn omx_paged_kv_cache_free(c: omx_paged_kv_cache) -> c_int;
    pub fn omx_paged_kv_cache_offset(c: omx_paged_kv_cache, offset: *mut c_int) -> c_int;
    pub fn omx_paged_kv_cache_lengths(c: omx_paged_kv_cache, lens: *mut i32) -> c_int;
    pub fn omx_paged_kv_cache_free_pages(c: omx_paged_kv_cache, n: *mut i64) -> c_int;
    pub fn omx_paged_kv_cache_reset(c: omx_paged_kv_cache, slot: c_int, s: omx_stream) -> c_int;
    pub fn omx_paged_kv_cache_release(c: omx_paged_kv_cache, slot: c_int, s: omx_stream) -> c_int;
    pub fn omx_paged_kv_cache_reserve(c: omx_paged_kv_cache, rows_ahead: c_int, s: omx_stream) -> c_int;
    pub fn omx_paged_kv_cache_sync_lengths(c: omx_paged_kv_cache, s: omx_stream) -> c_int;
    pub fn omx_paged_kv_cache_trim(c: omx_paged_kv_cache, n: c_int, s: omx_stream) -> c_int;
    pub fn omx_paged_kv_cache_update_and_fetch(c: omx_paged_kv_cache, keys: *const omx_array,
                                               values: *const omx_array, keys_out: *mut omx_array,
                                               values_out: *mut omx_array, s: omx_stream) -> c_int;
    pub fn omx_paged_kv_cache_append_slot(c: omx_paged_kv_cache, slot: c_int, keys: *const omx_array,
                                          values: *const omx_array, s: omx_stream) -> c_int;
    pub fn omx_paged_kv_cache_fetch(c: omx_paged_kv_cache, keys_out: *mut omx_array, values_out: *mut omx_array,
                                    s: omx_stream) -> c_int;
    pub fn omx_paged_kv_cache_pages(c: omx_paged_kv_cache, k_pool: *mut *mut c_void, v_pool: *mut *mut c_void,
                                    block_table: *mut *const i32, max_pages_per_seq: *mut c_int) -> c_int;
    pub fn omx_attn_decode_fused_paged(out: *const omx_array, q: *const omx_array, k_new: *const omx_array,
                                       v_new: *const omx_array, cache: omx_paged_kv_cache,
                                       q_norm_weight: *const omx_array, k_norm_weight: *const omx_array,
                                       norm_eps: f32, rope_dims: c_int, traditional: bool,
                                       base: omx_optional_float, rope_scale: f32, sm_scale: f32,
                                       s: omx_stream) -> c_int;
    pub fn omx_dit_rope(out: *const omx_array, x: *const omx_array, cos: *const omx_array,
                        sin: *const omx_array, s: omx_stream) -> c_int;
    pub fn omx_dit_joint_attention(out: *const omx_array, q: *const omx_array, k: *const omx_array,
                                   v: *const omx_array, scale: f32, add_mask: *const omx_array,
                                   s: omx_stream) -> c_int;
    pub fn omx_dit_attn_fused(out: *const omx_array, n_streams: c_int, q: *const *const omx_array,
                              k: *const *const omx_array, v: *const *const omx_array,
                              q_norm_weight: *const *const omx_array,
                              k_norm_weight: *const *const omx_array, norm_eps: f32,
                              cos: *const omx_array, sin: *const omx_array, scale: f32,
                              add_mask: *const omx_array, s: omx_stream) -> c_int;
    pub fn omx_last_kernel() -> *const c_char;
    pub fn omx_launch_count(reset: bool) -> i64;
    pub fn omx_force_kernel(name: *const c_char) -> c_int;
}
