"""Attention::forward for L > 1 as one composite call (omx_attn_prefill_fused) vs the oracle's op
chain (qwen3-mlx/src/model.rs:172-212): chunked prefill + decode, every mask spelling the crates
use, q_norm / k_norm on and off.  Outputs within tolerance, KV cache BIT-EXACT after every call."""
import numpy as np
import pytest
import torch

from conftest import assert_bits_equal, assert_close, load_oracle, load_pkg, n2f, randn, t2n, tdt

pytestmark = pytest.mark.gpu
omx = load_pkg()
orc = load_oracle()
DEV = "cuda"
Causal = omx.fast.ScaledDotProductAttentionMask.Causal


def _oracle_forward(oc, q, k, v, dtype, rope_t, scale, mask, qw=None, kw=None, eps=1e-6):
    if qw is not None:
        q = orc.rms_norm(q, qw, eps, dtype=dtype)
    if kw is not None:
        k = orc.rms_norm(k, kw, eps, dtype=dtype)
    off = oc.offset()
    if rope_t is not None:
        q = orc.rope(q, *rope_t, off, dtype=dtype)
        k = orc.rope(k, *rope_t, off, dtype=dtype)
    K, V = oc.update_and_fetch(k, v)
    return orc.sdpa(q, np.ascontiguousarray(K), np.ascontiguousarray(V), scale, mask, dtype=dtype)


def _call(gc, oc, B, Hq, Hkv, L, D, dtype, rope_t, mask_kind, seed, norms=None):
    # caller layout: projections are [B, L, H, D], viewed [B, H, L, D]
    q = randn((B, L, Hq, D), dtype, seed).transpose(1, 2)
    k = randn((B, L, Hkv, D), dtype, seed + 1).transpose(1, 2)
    v = randn((B, L, Hkv, D), dtype, seed + 2).transpose(1, 2)
    off = oc.offset()
    if mask_kind == "causal":
        gm, om = Causal, "causal"
    elif mask_kind == "array":  # create_attention_mask(h, cache, Some(true)) (utils.rs:156-188)
        gm = omx.create_causal_mask(L, off, device=DEV)
        om = orc.create_causal_mask(L, off)
    elif mask_kind == "none_multi":  # the callers' rule (model.rs:203-207): None && L > 1 -> Causal
        gm, om = None, "causal"
    else:
        gm = om = None
    rope = None if rope_t is None else omx.nn.Rope(*rope_t)
    qn = kn = qw = kw = None
    if norms:
        qw, kw = (1 + 0.1 * randn((D,), "f32", seed + 3)).to(tdt(dtype)), (1 + 0.1 * randn((D,), "f32", seed + 4)).to(tdt(dtype))
        qn, kn = omx.nn.RmsNorm(qw.to(DEV), 1e-6), omx.nn.RmsNorm(kw.to(DEV), 1e-6)
    got = omx.attn_prefill_fused(q.to(DEV), k.to(DEV), v.to(DEV), gc, rope, D ** -0.5, gm, q_norm=qn, k_norm=kn)
    want = _oracle_forward(oc, t2n(q, dtype), t2n(k, dtype), t2n(v, dtype), dtype, rope_t, D ** -0.5, om,
                           None if qw is None else t2n(qw, dtype), None if kw is None else t2n(kw, dtype))
    assert gc.offset() == oc.offset()
    assert_close(got.float().cpu().numpy(), n2f(want, dtype), dtype, f"prefill composite L={L} off={off} {mask_kind}")
    sk, sv = gc.state()
    assert_bits_equal(sk, oc.keys, dtype, "KV keys after the composite prefill")
    assert_bits_equal(sv, oc.values, dtype, "KV values after the composite prefill")
    return omx.last_kernel()


@pytest.mark.parametrize("norms", [False, True])
def test_chunked_prefill_then_decode_bf16(norms):
    B, Hq, Hkv, D, dtype = 2, 8, 2, 128, "bf16"
    rope_t = (D, False, 1e6, 1.0)
    gc, oc = omx.KVCache(), orc.KVCache()
    assert _call(gc, oc, B, Hq, Hkv, 300, D, dtype, rope_t, "causal", 10, norms) == "fmha_tcgen05"
    assert _call(gc, oc, B, Hq, Hkv, 77, D, dtype, rope_t, "array", 20, norms) == "fmha_tcgen05_arraymask"
    assert _call(gc, oc, B, Hq, Hkv, 200, D, dtype, rope_t, "causal", 30, norms) == "fmha_tcgen05"  # Lq < Lk: bottom-right
    assert _call(gc, oc, B, Hq, Hkv, 1, D, dtype, rope_t, "none", 40, norms).startswith("decode")
    assert gc.offset() == 578 and gc.state()[0].shape[2] == oc.keys.shape[2]


@pytest.mark.parametrize("dtype,D", [("f32", 128), ("f16", 64)])
def test_prefill_composite_generic_paths(dtype, D):
    gc, oc = omx.KVCache(), orc.KVCache()
    _call(gc, oc, 1, 4, 2, 50, D, dtype, (D, False, 10000.0, 1.0), "causal", 1, norms=True)
    _call(gc, oc, 1, 4, 2, 9, D, dtype, (D, False, 10000.0, 1.0), "array", 2, norms=True)


def test_qwen35_geometry_prefill_chunks_then_decode_steps():
    """Qwen3.5 full-attention layers (qwen3.5-35B-mlx/src/attention.rs): 16 q / 2 kv heads, head dim 256, rope on the first
    64 features, q / k norm.  Prefill chunks and single-token steps through the composites: the 512-byte-row prologue
    + mma.sync tiles; outputs vs the oracle chain, cache rows bit-exact."""
    B, Hq, Hkv, D, dtype = 2, 16, 2, 256, "bf16"
    rope_t = (64, False, 1e7, 1.0)
    gc, oc = omx.KVCache(), orc.KVCache()
    assert _call(gc, oc, B, Hq, Hkv, 150, D, dtype, rope_t, "causal", 10, True) == "sdpa_mma"
    assert _call(gc, oc, B, Hq, Hkv, 40, D, dtype, rope_t, "array", 20, True) == "sdpa_mma"
    for t in range(3):
        assert _call(gc, oc, B, Hq, Hkv, 1, D, dtype, rope_t, "none", 30 + 5 * t, True) == "sdpa_mma"
    assert gc.offset() == 193


def test_prefill_composite_partial_traditional_rope_and_no_rope():
    # glm4: traditional, partial rotary (glm4-mlx/src/model.rs:116-136)
    gc, oc = omx.KVCache(), orc.KVCache()
    _call(gc, oc, 1, 8, 2, 260, 128, "bf16", (64, True, 10000.0, 1.0), "array", 5)
    gc, oc = omx.KVCache(), orc.KVCache()
    _call(gc, oc, 1, 8, 2, 130, 128, "bf16", None, "causal", 6)


def test_prefill_composite_merged_head_output_layout():
    # N2: out given as the [B, L, Hq, D] storage viewed [B, Hq, L, D] -> the caller's transpose+reshape copy is gone
    B, Hq, Hkv, L, D = 2, 8, 2, 300, 128
    q, k, v = (randn((B, h, L, D), "bf16", s) for h, s in ((Hq, 1), (Hkv, 2), (Hkv, 3)))
    rope = omx.nn.Rope(D, False, 1e6, 1.0)
    ref = omx.attn_prefill_fused(q.to(DEV), k.to(DEV), v.to(DEV), omx.KVCache(), rope, D ** -0.5, Causal)
    merged = torch.empty((B, L, Hq * D), dtype=torch.bfloat16, device=DEV)
    omx.attn_prefill_fused(q.to(DEV), k.to(DEV), v.to(DEV), omx.KVCache(), rope, D ** -0.5, Causal,
                           out=merged.view(B, L, Hq, D).transpose(1, 2))
    assert torch.equal(merged, ref.transpose(1, 2).reshape(B, L, Hq * D))
    # and the decode kernel stores into the same kind of view
    c1, c2 = omx.KVCache(), omx.KVCache()
    for c in (c1, c2):
        c.update_and_fetch(k.to(DEV), v.to(DEV))
    q1, k1, v1 = (randn((B, h, 1, D), "bf16", s).to(DEV) for h, s in ((Hq, 4), (Hkv, 5), (Hkv, 6)))
    o_ref = omx.attn_decode_fused(q1, k1, v1, c1, rope, D ** -0.5)
    m1 = torch.empty((B, 1, Hq * D), dtype=torch.bfloat16, device=DEV)
    omx.attn_decode_fused(q1, k1, v1, c2, rope, D ** -0.5, out=m1.view(B, 1, Hq, D).transpose(1, 2))
    assert torch.equal(m1, o_ref.transpose(1, 2).reshape(B, 1, Hq * D))


def test_mask_none_with_several_new_tokens_is_causal():
    # Attention::forward(x, mask=None, cache) with L > 1 runs Causal in every LLM crate (qwen3-mlx/src/model.rs:203-207)
    B, Hq, Hkv, D, dtype = 1, 8, 2, 128, "bf16"
    gc, oc = omx.KVCache(), orc.KVCache()
    _call(gc, oc, B, Hq, Hkv, 40, D, dtype, (D, False, 1e6, 1.0), "causal", 5)
    _call(gc, oc, B, Hq, Hkv, 130, D, dtype, (D, False, 1e6, 1.0), "none_multi", 6)
