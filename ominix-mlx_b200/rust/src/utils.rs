//! mlx-rs-core/src/utils.rs for this path: SdpaMask, scaled_dot_product_attention, and the fused
//! decode step that replaces the rope/rope/update_and_fetch/sdpa sequence of Attention::forward.
use crate::array::{check, Array, Stream};
use crate::cache::{KVCache, KeyValueCache};
use crate::error::{Exception, Result};
use crate::fast::ScaledDotProductAttentionMask;
use crate::ffi;

/// Attention mask for scaled_dot_product_attention (utils.rs:105-116).
#[derive(Debug, Clone)]
pub enum SdpaMask<'a> {
    Causal,
    Array(&'a Array),
}

impl<'a> From<&'a Array> for SdpaMask<'a> {
    fn from(mask: &'a Array) -> Self {
        SdpaMask::Array(mask)
    }
}

/// Scaled dot-product attention (utils.rs:191-209): same signature, `_cache` unused there too.
pub fn scaled_dot_product_attention<'a, C>(
    queries: Array,
    keys: Array,
    values: Array,
    _cache: Option<C>,
    scale: f32,
    mask: Option<SdpaMask<'a>>,
) -> Result<Array>
where
    C: KeyValueCache,
{
    let sdpa_mask = match mask {
        Some(SdpaMask::Causal) => Some(ScaledDotProductAttentionMask::Causal),
        Some(SdpaMask::Array(m)) => Some(ScaledDotProductAttentionMask::Array(m)),
        None => None,
    };
    crate::fast::scaled_dot_product_attention(queries, keys, values, scale, sdpa_mask)
}

/// Rope parameters of `nn::Rope` (mlx-rs/src/nn/positional_encoding.rs:17-66).
#[derive(Debug, Clone, Copy)]
pub struct RopeParams {
    pub dimensions: i32,
    pub traditional: bool,
    pub base: f32,
    pub scale: f32,
}

/// The decode step of `Attention::forward` (qwen3-mlx/src/model.rs:186-212) in ONE launch:
/// `off = cache.offset(); q' = rope(q, off); k' = rope(k, off); cache.update_and_fetch(k', v);
/// sdpa(q', K, V, scale, None)`.  L must be 1.  Cache contents are bit-identical to the unfused calls.
pub fn attention_decode_fused(
    queries: &Array,
    keys: &Array,
    values: &Array,
    cache: &mut KVCache,
    rope: Option<RopeParams>,
    scale: f32,
    stream: Stream,
) -> Result<Array> {
    let qs = queries.shape();
    if qs.len() != 4 || qs[2] != 1 {
        return Err(Exception::custom("attention_decode_fused: queries must be [B, H, 1, D]"));
    }
    let out = Array::empty(&[qs[0], qs[1], 1, values.shape()[3]], queries.dtype())?;
    let (dims, trad, base, rscale) = match rope {
        Some(r) => (r.dimensions, r.traditional, ffi::omx_optional_float { value: r.base, has_value: true }, r.scale),
        None => (0, false, ffi::omx_optional_float::default(), 1.0),
    };
    check(unsafe {
        ffi::omx_attn_decode_fused(out.as_ptr(), queries.as_ptr(), keys.as_ptr(), values.as_ptr(), cache.raw(), dims,
                                   trad, base, rscale, std::ptr::null(), scale, std::ptr::null_mut(),
                                   std::ptr::null_mut(), stream.0)
    })?;
    let _ = cache.keepalive();
    Ok(out)
}

/// The same decode step over a `PagedKVCache`, ONE launch: per sequence the position is its own length (read by
/// the kernel from device memory), k' / v land in row `len % 64` of page `block_table[b][len / 64]`; released slots
/// are skipped.
pub fn attention_decode_fused_paged(
    queries: &Array,
    keys: &Array,
    values: &Array,
    cache: &mut crate::cache::PagedKVCache,
    rope: Option<RopeParams>,
    scale: f32,
    q_norm: Option<(&Array, f32)>,
    k_norm: Option<(&Array, f32)>,
    stream: Stream,
) -> Result<Array> {
    let qs = queries.shape();
    if qs.len() != 4 || qs[2] != 1 {
        return Err(Exception::custom("attention_decode_fused_paged: queries must be [B, H, 1, D]"));
    }
    let out = Array::empty(&[qs[0], qs[1], 1, values.shape()[3]], queries.dtype())?;
    let (dims, trad, base, rscale) = match rope {
        Some(r) => (r.dimensions, r.traditional, ffi::omx_optional_float { value: r.base, has_value: true }, r.scale),
        None => (0, false, ffi::omx_optional_float::default(), 1.0),
    };
    let eps = q_norm.map(|n| n.1).or(k_norm.map(|n| n.1)).unwrap_or(0.0);
    check(unsafe {
        ffi::omx_attn_decode_fused_paged(out.as_ptr(), queries.as_ptr(), keys.as_ptr(), values.as_ptr(), cache.raw(),
                                         q_norm.map_or(std::ptr::null(), |n| n.0.as_ptr()),
                                         k_norm.map_or(std::ptr::null(), |n| n.0.as_ptr()), eps, dims, trad, base,
                                         rscale, scale, stream.0)
    })?;
    let _ = cache.keepalive();
    Ok(out)
}

/// `attention_decode_fused` with the position read by the KERNEL from `position` (device `int32`, shared by all
/// layers): capturable with `cudaStreamBeginCapture` and replayable per token -- what `async_eval` pipelining gave
/// the reference's decode loop (qwen3-mlx/src/model.rs:798-844).  `out` is caller-owned (static under a graph);
/// the cache must have been pinned with `KVCache::prepare_graph`; the host offset is advanced separately with
/// `KVCache::advance(n)` after the replays.  `q_norm` / `k_norm`: optional `[D]` RMSNorm weights + eps.
#[allow(clippy::too_many_arguments)]
pub fn attention_decode_fused_dynamic(
    out: &Array,
    queries: &Array,
    keys: &Array,
    values: &Array,
    cache: &KVCache,
    rope: Option<RopeParams>,
    scale: f32,
    position: *const i32,
    q_norm: Option<(&Array, f32)>,
    k_norm: Option<(&Array, f32)>,
    stream: Stream,
) -> Result<()> {
    let (dims, trad, base, rscale) = match rope {
        Some(r) => (r.dimensions, r.traditional, ffi::omx_optional_float { value: r.base, has_value: true }, r.scale),
        None => (0, false, ffi::omx_optional_float::default(), 1.0),
    };
    let eps = q_norm.map(|n| n.1).or(k_norm.map(|n| n.1)).unwrap_or(0.0);
    check(unsafe {
        ffi::omx_attn_decode_fused_dynamic(out.as_ptr(), queries.as_ptr(), keys.as_ptr(), values.as_ptr(), cache.raw(),
                                           q_norm.map_or(std::ptr::null(), |n| n.0.as_ptr()),
                                           k_norm.map_or(std::ptr::null(), |n| n.0.as_ptr()), eps, dims, trad, base,
                                           rscale, scale, position, stream.0)
    })
}

/// `*counter += delta` on the stream: the per-token position bump of a graph-captured decode loop.
pub fn device_counter_add(counter: *mut i32, delta: i32, stream: Stream) -> Result<()> {
    check(unsafe { ffi::omx_device_counter_add(counter, delta, stream.0) })
}
