"""Mirror of mlx_rs::fast for the attention path (mlx-rs/src/fast.rs).

Same names, argument order and meaning as the Rust functions; arrays are torch CUDA tensors in
the reference's [B, H, L, D] layout (strided views welcome).  Errors surface as
`Exception_` carrying the library message, like `Exception{what}` in mlx-rs.
"""
import enum

import torch

from . import _lib
from .array import desc, ref, stream_ptr


class ScaledDotProductAttentionMask(enum.Enum):
    """mlx-rs/src/fast.rs:53-62.  `Array(t)` / `Arrays([t, ...])` are passed as the tensor / list."""
    Causal = "causal"


def _opt_float(x):
    o = _lib.OmxOptionalFloat()
    o.has_value = x is not None
    o.value = float(x) if x is not None else 0.0
    return o


def rms_norm(x, weight, eps, stream=None, out=None):
    """fast::rms_norm (mlx-rs/src/fast.rs:163-180): normalisation over the last axis, `weight` [D] in x's
    dtype (None for no scaling, as mlx_fast_rms_norm allows)."""
    if out is None:
        out = torch.empty(x.shape, dtype=x.dtype, device=x.device)
    xd, od, wd = desc(x), desc(out), desc(weight)
    _lib.check(_lib.lib().omx_fast_rms_norm(ref(od), ref(xd), ref(wd), float(eps), stream_ptr(stream)))
    return out


def rope(array, dimensions, traditional, base, scale, offset, freqs=None, stream=None, out=None):
    """fast::rope (mlx-rs/src/fast.rs:15-46).  `offset` may be an int or an int32 CUDA scalar tensor
    (mlx_fast_rope_dynamic); in the latter case pass `max_position` via `rope_dynamic`."""
    if isinstance(offset, torch.Tensor):
        raise _lib.Exception_("use rope_dynamic(...) for a device-resident offset")
    if out is None:
        out = torch.empty(array.shape, dtype=array.dtype, device=array.device)
    x, o, f = desc(array), desc(out), desc(freqs)
    _lib.check(_lib.lib().omx_fast_rope(ref(o), ref(x), int(dimensions), bool(traditional), _opt_float(base),
                                        float(scale), int(offset), ref(f), stream_ptr(stream)))
    return out


def rope_dynamic(array, dimensions, traditional, base, scale, offset, max_position, freqs=None, stream=None,
                 out=None):
    """mlx_fast_rope_dynamic (mlx-c/mlx/c/fast.h:179-188): offset is an int32 scalar in device memory."""
    if out is None:
        out = torch.empty(array.shape, dtype=array.dtype, device=array.device)
    x, o, f, off = desc(array), desc(out), desc(freqs), desc(offset)
    _lib.check(_lib.lib().omx_fast_rope_dynamic(ref(o), ref(x), int(dimensions), bool(traditional),
                                                _opt_float(base), float(scale), ref(off), int(max_position),
                                                ref(f), stream_ptr(stream)))
    return out


def _mode_and_mask(mask):
    """ScaledDotProductAttentionMask::as_mode_and_mask_ptr (mlx-rs/src/fast.rs:88-108)."""
    if mask is None:
        return b"", None
    if mask is ScaledDotProductAttentionMask.Causal or (isinstance(mask, str) and mask == "causal"):
        return b"causal", None
    if isinstance(mask, (list, tuple)):  # Arrays: the new API only uses the first one
        return b"", (mask[0] if len(mask) else None)
    if isinstance(mask, torch.Tensor):
        return b"", mask
    raise _lib.Exception_(f"unsupported mask {mask!r}")


def scaled_dot_product_attention(queries, keys, values, scale, mask=None, stream=None, out=None):
    """fast::scaled_dot_product_attention (mlx-rs/src/fast.rs:110-151).
    O = softmax(scale * Q K^T + mask) V, GQA without pre-tiling, f32 softmax; output [B,Hq,Lq,Dv]."""
    mode, m = _mode_and_mask(mask)
    if out is None:
        if queries.dim() == 4 and values.dim() == 4:
            shape = (queries.shape[0], queries.shape[1], queries.shape[2], values.shape[3])
        else:  # rank errors are reported by the library, with the reference's message
            shape = tuple(queries.shape)
        out = torch.empty(shape, dtype=queries.dtype, device=queries.device)
    q, k, v, o, md = desc(queries), desc(keys), desc(values), desc(out), desc(m)
    _lib.check(_lib.lib().omx_fast_scaled_dot_product_attention(ref(o), ref(q), ref(k), ref(v), float(scale),
                                                                mode, ref(md), None, stream_ptr(stream)))
    return out
