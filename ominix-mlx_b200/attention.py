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


def attn_decode_fused_dynamic(q, k_new, v_new, cache, rope, sm_scale, position, stream=None, out=None,
                              q_norm=None, k_norm=None):
    """attn_decode_fused with the position read by the kernel from `position` (int32 CUDA tensor, one
    element, shared by all layers) -- capturable into a CUDA graph and replayable for every token.
    The cache must have been pinned with cache.prepare_graph(max_rows, Hq); the host offset is NOT
    advanced here (cache.advance(n) after the step / replays)."""
    if out is None:
        out = torch.empty((q.shape[0], q.shape[1], 1, v_new.shape[3]), dtype=q.dtype, device=q.device)
    if position.dtype != torch.int32 or not position.is_cuda or position.numel() != 1:
        raise _lib.Exception_("position must be a one-element int32 CUDA tensor")
    base = _lib.OmxOptionalFloat()
    base.has_value = rope is not None
    base.value = rope.base if rope is not None else 0.0
    eps = (q_norm or k_norm).eps if (q_norm is not None or k_norm is not None) else 0.0
    if q_norm is not None and k_norm is not None and q_norm.eps != k_norm.eps:
        raise _lib.Exception_("q_norm and k_norm must share one eps in the fused step")
    qd, kd, vd, od = desc(q), desc(k_new), desc(v_new), desc(out)
    qw = desc(q_norm.weight) if q_norm is not None else None
    kw = desc(k_norm.weight) if k_norm is not None else None
    _lib.check(_lib.lib().omx_attn_decode_fused_dynamic(
        ref(od), ref(qd), ref(kd), ref(vd), cache.handle, ref(qw), ref(kw), float(eps),
        int(rope.dimensions if rope is not None else 0), bool(rope.traditional) if rope is not None else False,
        base, float(rope.scale) if rope is not None else 1.0, float(sm_scale), position.data_ptr(),
        stream_ptr(stream)))
    return out


def attn_decode_fused_paged(q, k_new, v_new, cache, rope, sm_scale, stream=None, out=None, q_norm=None, k_norm=None):
    """attn_decode_fused over a PagedKVCache, ONE launch: per sequence b the position is its own length
    (read by the kernel from device memory), k' / v_new land in row len % 64 of page block_table[b][len / 64],
    K/V tiles are fetched page by page.  Released slots are skipped (their output rows are left untouched)."""
    if out is None:
        out = torch.empty((q.shape[0], q.shape[1], 1, v_new.shape[3]), dtype=q.dtype, device=q.device)
    base = _lib.OmxOptionalFloat()
    base.has_value = rope is not None
    base.value = rope.base if rope is not None else 0.0
    eps = (q_norm or k_norm).eps if (q_norm is not None or k_norm is not None) else 0.0
    if q_norm is not None and k_norm is not None and q_norm.eps != k_norm.eps:
        raise _lib.Exception_("q_norm and k_norm must share one eps in the fused step")
    qd, kd, vd, od = desc(q), desc(k_new), desc(v_new), desc(out)
    qw = desc(q_norm.weight) if q_norm is not None else None
    kw = desc(k_norm.weight) if k_norm is not None else None
    _lib.check(_lib.lib().omx_attn_decode_fused_paged(
        ref(od), ref(qd), ref(kd), ref(vd), cache.handle, ref(qw), ref(kw), float(eps),
        int(rope.dimensions if rope is not None else 0), bool(rope.traditional) if rope is not None else False,
        base, float(rope.scale) if rope is not None else 1.0, float(sm_scale), stream_ptr(stream)))
    return out


def device_counter_add(counter, delta=1, stream=None):
    """counter (int32 CUDA tensor) += delta on the stream: the per-token position bump of a graph loop."""
    _lib.check(_lib.lib().omx_device_counter_add(counter.data_ptr(), int(delta), stream_ptr(stream)))


class DecodeLoopGraph:
    """One decode step of an n-layer model's attention -- per layer the one-launch fused step
    (q_norm/k_norm, rope, KV append, GQA attention), then the position bump -- captured ONCE into a CUDA
    graph and replayed per token: what MLX's lazy graph + async_eval gave the reference's decode loop
    (qwen3-mlx/src/model.rs:798-844), without re-tracing.

    q / k_new / v_new / out are STATIC per-layer buffers (lists of tensors): the projections of the
    surrounding model write into them before each replay, as with any CUDA graph.  `caches`: one
    prefilled KVCache per layer, all at the same offset."""

    def __init__(self, q, k_new, v_new, caches, rope, sm_scale, max_rows, q_norm=None, k_norm=None, out=None):
        n = len(caches)
        self.caches, self.q, self.k_new, self.v_new = caches, q, k_new, v_new
        dev = q[0].device
        off = caches[0].offset()
        if any(c.offset() != off for c in caches):
            raise _lib.Exception_("all layer caches must be at the same offset")
        self.max_rows = int(max_rows)
        self.out = out or [torch.empty((t.shape[0], t.shape[1], 1, v.shape[3]), dtype=t.dtype, device=dev)
                           for t, v in zip(q, v_new)]
        self.position = torch.full((1,), off, dtype=torch.int32, device=dev)
        qn = q_norm if isinstance(q_norm, (list, tuple)) else [q_norm] * n
        kn = k_norm if isinstance(k_norm, (list, tuple)) else [k_norm] * n
        for c, t in zip(caches, q):
            c.prepare_graph(max_rows, t.shape[1])

        def step(stream=None):
            for i in range(n):
                attn_decode_fused_dynamic(q[i], k_new[i], v_new[i], caches[i], rope, sm_scale, self.position,
                                          stream=stream, out=self.out[i], q_norm=qn[i], k_norm=kn[i])
            device_counter_add(self.position, 1, stream=stream)
        # eager warm-up on a side stream builds the rope table and proves the launch configuration; its
        # effects are rolled back (same row is rewritten by the first replay, position restored)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            step()
            device_counter_add(self.position, -1)
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            step()
        self.launches_per_step = n + 1

    def step(self):
        """Replay one token step (asynchronous) and advance the host-side offsets."""
        if self.caches[0].offset() + 1 > self.max_rows:
            raise _lib.Exception_(f"decode loop graph was pinned for {self.max_rows} rows")
        self.graph.replay()
        for c in self.caches:
            c.advance(1)
        return self.out


def attn_prefill_fused(q, k_new, v_new, cache, rope, sm_scale, mask=None, stream=None, out=None, fetch=False,
                       q_norm=None, k_norm=None):
    """Attention::forward for L >= 1 new tokens (model.rs:172-212) with the minimum of memory passes:
    k' = rope(k_norm(k_new), off) lands directly in the cache rows, v_new is copied into its rows, then
    sdpa(rope(q_norm(q), off), K, V, sm_scale, mask).  `mask`: None, ScaledDotProductAttentionMask.Causal /
    "causal", or a bool / additive tensor (what create_attention_mask returns).  The callers' rule applies
    (model.rs:203-207): `None` with L > 1 means Causal; pass mask=False for an unmasked multi-token call."""
    if out is None:
        out = torch.empty((q.shape[0], q.shape[1], q.shape[2], v_new.shape[3]), dtype=q.dtype, device=q.device)
    if mask is None and q.shape[2] > 1:
        mask = fast.ScaledDotProductAttentionMask.Causal
    elif mask is False:
        mask = None
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
