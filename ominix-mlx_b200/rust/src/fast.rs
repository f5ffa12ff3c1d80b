//! `mlx_rs::fast::{rope, scaled_dot_product_attention}` with the reference's signatures
//! (mlx-rs/src/fast.rs:15-46, 53-108, 110-151), executed eagerly by libomx_attn on a CUDA stream.
use std::ffi::CString;

use crate::array::{check, Array, Stream};
use crate::error::{Exception, Result};
use crate::ffi;

/// fast.rs:15-46.  `base` and `freqs` are mutually exclusive, exactly as in the reference
/// (`fast.rs:25-29,40-42`); the rotation keeps the MLX CPU backend's per-op rounding.
#[allow(clippy::too_many_arguments)]
pub fn rope_device<'a>(
    array: impl AsRef<Array>,
    dimensions: i32,
    traditional: bool,
    base: impl Into<Option<f32>>,
    scale: f32,
    offset: i32,
    freqs: impl Into<Option<&'a Array>>,
    stream: Stream,
) -> Result<Array> {
    let x = array.as_ref();
    let base = match base.into() {
        Some(value) => ffi::omx_optional_float { value, has_value: true },
        None => ffi::omx_optional_float { value: 0.0, has_value: false },
    };
    let freqs = freqs.into();
    let out = Array::empty(x.shape(), x.dtype())?;
    check(unsafe {
        ffi::omx_fast_rope(out.as_ptr(), x.as_ptr(), dimensions, traditional, base, scale, offset,
                           freqs.map_or(std::ptr::null(), |f| f.as_ptr()), stream.0)
    })?;
    Ok(out)
}

/// The `#[default_device]` twin of [`rope_device`] (mlx-internal-macros/src/lib.rs:83-94).
pub fn rope<'a>(
    array: impl AsRef<Array>,
    dimensions: i32,
    traditional: bool,
    base: impl Into<Option<f32>>,
    scale: f32,
    offset: i32,
    freqs: impl Into<Option<&'a Array>>,
) -> Result<Array> {
    rope_device(array, dimensions, traditional, base, scale, offset, freqs, Stream::default())
}

/// fast.rs:163-180: RMS normalisation over the last axis (`weight`: `[D]` in `x`'s dtype).
pub fn rms_norm_device(x: impl AsRef<Array>, weight: impl AsRef<Array>, eps: f32, stream: Stream) -> Result<Array> {
    let x = x.as_ref();
    let out = Array::empty(x.shape(), x.dtype())?;
    check(unsafe { ffi::omx_fast_rms_norm(out.as_ptr(), x.as_ptr(), weight.as_ref().as_ptr(), eps, stream.0) })?;
    Ok(out)
}

pub fn rms_norm(x: impl AsRef<Array>, weight: impl AsRef<Array>, eps: f32) -> Result<Array> {
    rms_norm_device(x, weight, eps, Stream::default())
}

/// fast.rs:53-62.
#[derive(Debug, Clone)]
pub enum ScaledDotProductAttentionMask<'a> {
    Array(&'a Array),
    Arrays(&'a [Array]),
    Causal,
}

impl<'a> From<&'a Array> for ScaledDotProductAttentionMask<'a> {
    fn from(mask: &'a Array) -> Self {
        ScaledDotProductAttentionMask::Array(mask)
    }
}
impl<'a> From<&'a [Array]> for ScaledDotProductAttentionMask<'a> {
    fn from(masks: &'a [Array]) -> Self {
        ScaledDotProductAttentionMask::Arrays(masks)
    }
}

impl ScaledDotProductAttentionMask<'_> {
    /// fast.rs:88-108: mode string + optional mask pointer; `Arrays` only uses the first entry.
    fn as_mode_and_mask(&self) -> (&'static str, *const ffi::omx_array) {
        match self {
            ScaledDotProductAttentionMask::Array(m) => ("", m.as_ptr()),
            ScaledDotProductAttentionMask::Arrays(ms) => ("", ms.first().map_or(std::ptr::null(), |m| m.as_ptr())),
            ScaledDotProductAttentionMask::Causal => ("causal", std::ptr::null()),
        }
    }
}

/// fast.rs:110-151.  O = softmax(scale * Q K^T + mask) V; GQA without pre-tiling K/V; softmax and
/// accumulation in f32; output `[B, Hq, Lq, Dv]` in the input dtype.
pub fn scaled_dot_product_attention_device<'a>(
    queries: impl AsRef<Array>,
    keys: impl AsRef<Array>,
    values: impl AsRef<Array>,
    scale: f32,
    mask: impl Into<Option<ScaledDotProductAttentionMask<'a>>>,
    stream: Stream,
) -> Result<Array> {
    let (q, k, v) = (queries.as_ref(), keys.as_ref(), values.as_ref());
    if q.ndim() != 4 || v.ndim() != 4 {
        return Err(Exception::custom(format!(
            "[scaled_dot_product_attention] input with shape of {} dims is not supported; expected [B, N, T, D]",
            q.ndim()
        )));
    }
    let (qs, vs) = (q.shape(), v.shape());
    let out = Array::empty(&[qs[0], qs[1], qs[2], vs[3]], q.dtype())?;
    let mask = mask.into();
    let (mode, mask_ptr) = mask.as_ref().map_or(("", std::ptr::null()), |m| m.as_mode_and_mask());
    let mode = CString::new(mode).expect("static mode string");
    check(unsafe {
        ffi::omx_fast_scaled_dot_product_attention(out.as_ptr(), q.as_ptr(), k.as_ptr(), v.as_ptr(), scale,
                                                   mode.as_ptr(), mask_ptr, std::ptr::null(), stream.0)
    })?;
    Ok(out)
}

pub fn scaled_dot_product_attention<'a>(
    queries: impl AsRef<Array>,
    keys: impl AsRef<Array>,
    values: impl AsRef<Array>,
    scale: f32,
    mask: impl Into<Option<ScaledDotProductAttentionMask<'a>>>,
) -> Result<Array> {
    scaled_dot_product_attention_device(queries, keys, values, scale, mask, Stream::default())
}
