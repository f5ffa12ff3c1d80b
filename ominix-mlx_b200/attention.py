"""The composite decode step of the reference's Attention::forward
(qwen3-mlx/src/model.rs:186-212) as ONE kernel launch, plus its unfused spelling."""
import torch

from . import _lib, fast
from .array import desc, ref, stream_ptr, view


def attn_decode_fused(q, k_new, v_new, cache, rope, sm_scale, stream=None, out=None, fetch=False,
                      q_norm=None, k_norm=None):
    """off = cache.offset(); q' = rope(q, off); k' = rope(k_new, off);
    (K, V) = cache.update_and_fetch(k', v_new); out = sdpa(q', K, V, sm_scale)   -- L == 1 only.
    rope: an nn.Rope (or None to skip the rotation).  q_norm / k_norm: nn.RmsNorm applied per head before
    the rotation (Qwen3: model.rs:172-181), fused into the same launch.
    Returns out, or (out, K, V) with fetch=True."""
    if out is None:
        out = torch.empty((q.shape[0], q.shape[1], 1, v_new.shape[3]), dtype=q.dtype, device=q.device)
    dims = rope.dimensions if rope is not None else 0
    base = _lib.OmxOptionalFloat()
    base.has_value = rope is not None
    base.value = rope.base if rope is not None else 0.0
    ko, vo = _lib.OmxArray(), _lib.OmxArray()
    qd, kd, vd, od = desc(q), desc(k_new), desc(v_new), desc(out)
    trad = bool(rope.traditional) if rope is not None else False
    rscale = float(rope.scale) if rope is not None else 1.0
    if q_norm is None and k_norm is None:
        _lib.check(_lib.lib().omx_attn_decode_fused(
            ref(od), ref(qd), ref(kd), ref(vd), cache.handle, int(dims), trad, base, rscale, None,
            float(sm_scale), ref(ko), ref(vo), stream_ptr(stream)))
    else:
        eps = (q_norm or k_norm).eps
        if q_norm is not None and k_norm is not None and q_norm.eps != k_norm.eps:
            raise _lib.Exception_("q_norm and k_norm must share one eps in the fused step")
        qw = desc(q_norm.weight) if q_norm is not None else None
        kw = desc(k_norm.weight) if k_norm is not None else None
        _lib.check(_lib.lib().omx_attn_decode_fused_norm(
            ref(od), ref(qd), ref(kd), ref(vd), cache.handle, ref(qw), ref(kw), float(eps), int(dims), trad,
            base, rscale, None, float(sm_scale), ref(ko), ref(vo), stream_ptr(stream)))
    if fetch:
        return out, view(ko, cache, q.device), view(vo, cache, q.device)
    return out


def attn_prefill_fused(q, k_new, v_new, cache, rope, sm_scale, mask=None, stream=None, out=None, fetch=False,
                       q_norm=None, k_norm=None):
    """Attention::forward for L >= 1 new tokens (model.rs:172-212) with the minimum of memory passes:
    k' = rope(k_norm(k_new), off) lands directly in the cache rows, v_new is copied into its rows, then
    sdpa(rope(q_norm(q), off), K, V, sm_scale, mask).  `mask`: None, ScaledDotProductAttentionMask.Causal /
    "causal", or a bool / additive tensor (what create_attention_mask returns)."""
    if out is None:
        out = torch.empty((q.shape[0], q.shape[1], q.shape[2], v_new.shape[3]), dtype=q.dtype, device=q.device)
    mode, m = fast._mode_and_mask(mask)
    base = _lib.OmxOptionalFloat()
    base.has_value = rope is not None
    base.value = rope.base if rope is not None else 0.0
    eps = (q_norm or k_norm).eps if (q_norm is not None or k_norm is not None) else 0.0
    ko, vo = _lib.OmxArray(), _lib.OmxArray()
    qd, kd, vd, od, md = desc(q), desc(k_new), desc(v_new), desc(out), desc(m)
    qw = desc(q_norm.weight) if q_norm is not None else None
    kw = desc(k_norm.weight) if k_norm is not None else None
    _lib.check(_lib.lib().omx_attn_prefill_fused(
        ref(od), ref(qd), ref(kd), ref(vd), cache.handle, ref(qw), ref(kw), float(eps),
        int(rope.dimensions if rope is not None else 0), bool(rope.traditional) if rope is not None else False,
        base, float(rope.scale) if rope is not None else 1.0, None, float(sm_scale), mode, ref(md), ref(ko),
        ref(vo), stream_ptr(stream)))
    if fetch:
        return out, view(ko, cache, q.device), view(vo, cache, q.device)
    return out


def attn_decode_unfused(q, k_new, v_new, cache, rope, sm_scale, stream=None, q_norm=None, k_norm=None, mask=None):
    """The reference's op sequence, one library call per op (model.rs:172-212)."""
    if q_norm is not None:
        q = q_norm.forward(q, stream)
    if k_norm is not None:
        k_new = k_norm.forward(k_new, stream)
    off = cache.offset()
    if rope is not None:
        q = rope.forward(q, off, stream)
        k_new = rope.forward(k_new, off, stream)
    keys, values = cache.update_and_fetch(k_new, v_new, stream)
    if mask is None and q.shape[2] > 1:  # the callers' rule: None && L > 1 -> Causal (model.rs:203-207)
        mask = fast.ScaledDotProductAttentionMask.Causal
    return fast.scaled_dot_product_attention(q, keys, values, sm_scale, mask, stream)
