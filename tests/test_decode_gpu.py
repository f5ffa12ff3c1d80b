"""The fused decode step (rope + KV append + split-K GQA attention in one launch) vs the
oracle's op-by-op chain: attention within tolerance, KV cache contents BIT-EXACT."""
import numpy as np
import pytest
import torch

from conftest import assert_bits_equal, assert_close, load_oracle, load_pkg, n2f, randn, t2n

pytestmark = pytest.mark.gpu
omx = load_pkg()
orc = load_oracle()
DEV = "cuda"


def _oracle_step(ocache, q, k_new, v_new, dtype, rope, scale):
    """Attention::forward decode step, qwen3-mlx/src/model.rs:186-212."""
    off = ocache.offset()
    if rope is not None:
        dims, trad, base, rs = rope
        q = orc.rope(q, dims, trad, base, rs, off, dtype=dtype)
        k_new = orc.rope(k_new, dims, trad, base, rs, off, dtype=dtype)
    K, V = ocache.update_and_fetch(k_new, v_new)
    return orc.sdpa(q, np.ascontiguousarray(K), np.ascontiguousarray(V), scale, None, dtype=dtype)


def _prefill(B, Hkv, S, D, dtype, seed):
    gc, oc = omx.KVCache(), orc.KVCache()
    if S > 0:
        k = randn((B, Hkv, S, D), dtype, seed)
        v = randn((B, Hkv, S, D), dtype, seed + 1)
        gc.update_and_fetch(k.to(DEV), v.to(DEV))
        oc.update_and_fetch(t2n(k, dtype), t2n(v, dtype))
    return gc, oc


def _step(gc, oc, B, Hq, Hkv, D, dtype, rope_t, seed, strided=True):
    if strided:  # projections come out as [B, L, H, D] and are viewed [B, H, L, D]
        q = randn((B, 1, Hq, D), dtype, seed).transpose(1, 2)
        k = randn((B, 1, Hkv, D), dtype, seed + 1).transpose(1, 2)
        v = randn((B, 1, Hkv, D), dtype, seed + 2).transpose(1, 2)
    else:
        q = randn((B, Hq, 1, D), dtype, seed)
        k = randn((B, Hkv, 1, D), dtype, seed + 1)
        v = randn((B, Hkv, 1, D), dtype, seed + 2)
    rope = None if rope_t is None else omx.nn.Rope(rope_t[0], rope_t[1], rope_t[2], rope_t[3])
    scale = D ** -0.5
    got = omx.attn_decode_fused(q.to(DEV), k.to(DEV), v.to(DEV), gc, rope, scale)
    want = _oracle_step(oc, t2n(q, dtype), t2n(k, dtype), t2n(v, dtype), dtype, rope_t, scale)
    assert gc.offset() == oc.offset()
    assert_close(got.float().cpu().numpy(), n2f(want, dtype), dtype, f"fused decode off={oc.offset()}")
    sk, sv = gc.state()
    assert_bits_equal(sk, oc.keys, dtype, "KV cache keys after fused append")
    assert_bits_equal(sv, oc.values, dtype, "KV cache values after fused append")
    return got


ROPE = (128, False, 1e6, 1.0)


def test_config_c1_qwen3_0p6b_fp32_decode():
    # C1: 16 q / 8 kv heads, D128, fp32, B1, ctx 2048 (append at offset 2047)
    gc, oc = _prefill(1, 8, 2047, 128, "f32", 10)
    _step(gc, oc, 1, 16, 8, 128, "f32", ROPE, 20)
    assert omx.last_kernel() == "decode_simt"
    assert gc.offset() == 2048 and gc.state()[0].shape[2] == 2048


def test_config_c2_shape_reduced_batch_bf16():
    # C2 geometry (32 q / 8 kv, D128, bf16, ctx 8192) at B=4
    gc, oc = _prefill(4, 8, 8191, 128, "bf16", 11)
    _step(gc, oc, 4, 32, 8, 128, "bf16", ROPE, 21)
    assert omx.last_kernel() == "decode_hmma_tma"
    assert gc.offset() == 8192


def test_config_c5_shape_single_sequence_bf16():
    # C5: Mixtral-shape B1, ctx 32768 (one rank's view when kv-head sharded: Hkv=1, Hq=4)
    gc, oc = _prefill(1, 1, 32767, 128, "bf16", 12)
    _step(gc, oc, 1, 4, 1, 128, "bf16", (128, False, 1e6, 1.0), 22)
    assert gc.offset() == 32768


@pytest.mark.parametrize("dtype", ["bf16", "f16", "f32"])
@pytest.mark.parametrize("G", [1, 2, 4, 8, 16])
def test_group_sizes(dtype, G):
    Hkv = 2
    gc, oc = _prefill(2, Hkv, 300, 128, dtype, 13 + G)
    _step(gc, oc, 2, Hkv * G, Hkv, 128, dtype, ROPE, 23 + G)


@pytest.mark.parametrize("dtype", ["bf16", "f32"])
@pytest.mark.parametrize("rope_t", [None, (128, True, 1e6, 1.0), (64, True, 10000.0, 1.0), (64, False, 10000.0, 0.5)])
def test_rope_variants_glm4_mixtral(dtype, rope_t):
    # glm4: partial + traditional (glm4-mlx/src/model.rs:117,133-136); mixtral: traditional flag
    gc, oc = _prefill(2, 2, 130, 128, dtype, 14)
    _step(gc, oc, 2, 8, 2, 128, dtype, rope_t, 24)


@pytest.mark.parametrize("dtype", ["bf16", "f32"])
def test_first_token_and_growth_across_step_boundary(dtype):
    # empty cache (n_mem == 0), then decode across the 256-row growth boundary
    gc, oc = _prefill(1, 2, 0, 128, dtype, 15)
    for i in range(3):
        _step(gc, oc, 1, 8, 2, 128, dtype, ROPE, 100 + 3 * i)
    gc, oc = _prefill(1, 2, 254, 128, dtype, 16)
    for i in range(4):
        _step(gc, oc, 1, 8, 2, 128, dtype, ROPE, 200 + 3 * i)
    assert gc.offset() == 258 and gc.state()[0].shape[2] == 512


@pytest.mark.parametrize("S", [1, 63, 64, 65, 127, 128, 1000, 4097])
def test_ragged_lengths_bf16(S):
    gc, oc = _prefill(3, 2, S, 128, "bf16", 17 + S)
    _step(gc, oc, 3, 8, 2, 128, "bf16", ROPE, 27 + S)


def test_other_head_dims_go_through_simt_or_generic():
    for D in (64, 256):
        gc, oc = _prefill(2, 2, 200, D, "bf16", 18)
        _step(gc, oc, 2, 4, 2, D, "bf16", (D, False, 10000.0, 1.0), 28)
    gc, oc = _prefill(2, 2, 200, 80, "f32", 19)  # D=80: unfused composition, same results
    _step(gc, oc, 2, 4, 2, 80, "f32", (80, False, 10000.0, 1.0), 29)
    assert omx.last_kernel() == "sdpa_generic"


def test_fused_equals_unfused_library_calls():
    B, Hq, Hkv, S, D = 4, 32, 8, 1500, 128
    k = randn((B, Hkv, S, D), "bf16", 1).to(DEV)
    v = randn((B, Hkv, S, D), "bf16", 2).to(DEV)
    c1, c2 = omx.KVCache(), omx.KVCache()
    c1.update_and_fetch(k, v)
    c2.update_and_fetch(k, v)
    rope = omx.nn.Rope(128, False, 1e6, 1.0)
    q = randn((B, Hq, 1, D), "bf16", 3).to(DEV)
    kn = randn((B, Hkv, 1, D), "bf16", 4).to(DEV)
    vn = randn((B, Hkv, 1, D), "bf16", 5).to(DEV)
    a = omx.attn_decode_fused(q, kn, vn, c1, rope, D ** -0.5)
    n0 = omx.launch_count(reset=True)
    a2 = omx.attn_decode_fused(q, kn, vn, c1, rope, D ** -0.5)
    assert omx.launch_count() == 1, "fused decode must be ONE kernel launch"
    b = omx.attn_decode_unfused(q, kn, vn, c2, rope, D ** -0.5)
    assert (a.float() - b.float()).abs().max().item() <= 1e-2
    assert torch.equal(c1.state()[0][:, :, :S + 1], c2.state()[0][:, :, :S + 1])
    assert torch.equal(c1.state()[1][:, :, :S + 1], c2.state()[1][:, :, :S + 1])
    assert a2.shape == a.shape and n0 >= 1


def test_full_size_c2_against_oracle_and_properties():
    # BASELINE C2 at full size: B64, 32q/8kv, D128, bf16, ctx 8192 (2 GiB of KV)
    B, Hq, Hkv, S, D = 64, 32, 8, 8192, 128
    g = torch.Generator(device=DEV).manual_seed(1236)
    k = torch.randn((B, Hkv, S - 1, D), generator=g, device=DEV, dtype=torch.float32).bfloat16()
    v = torch.randn((B, Hkv, S - 1, D), generator=g, device=DEV, dtype=torch.float32).bfloat16()
    c = omx.KVCache()
    c.update_and_fetch(k, v)
    q = torch.randn((B, Hq, 1, D), generator=g, device=DEV, dtype=torch.float32).bfloat16()
    kn = torch.randn((B, Hkv, 1, D), generator=g, device=DEV, dtype=torch.float32).bfloat16()
    vn = torch.randn((B, Hkv, 1, D), generator=g, device=DEV, dtype=torch.float32).bfloat16()
    rope = omx.nn.Rope(128, False, 1e6, 1.0)
    out, K, V = omx.attn_decode_fused(q, kn, vn, c, rope, D ** -0.5, fetch=True)
    assert omx.last_kernel() == "decode_hmma_tma" and c.offset() == S
    # (1) the whole thing against the oracle (a few seconds of CPU)
    qo = orc.rope(t2n(q, "bf16"), 128, False, 1e6, 1.0, S - 1, dtype="bf16")
    ko = orc.rope(t2n(kn, "bf16"), 128, False, 1e6, 1.0, S - 1, dtype="bf16")
    assert_bits_equal(K[:, :, S - 1:S], ko, "bf16", "appended key row")
    assert_bits_equal(V[:, :, S - 1:S], t2n(vn, "bf16"), "bf16", "appended value row")
    assert torch.equal(K[:, :, :S - 1], k) and torch.equal(V[:, :, :S - 1], v)
    want = orc.sdpa(qo, t2n(K, "bf16"), t2n(V, "bf16"), D ** -0.5, None, dtype="bf16")
    assert_close(out.float().cpu().numpy(), n2f(want, "bf16"), "bf16", "C2 full size")
    # (2) size-independent properties: key-order invariance and linearity in V
    qr = omx.fast.rope(q, 128, False, 1e6, 1.0, S - 1)
    perm = torch.randperm(S, device=DEV)
    o1 = omx.fast.scaled_dot_product_attention(qr, K, V, D ** -0.5)
    o2 = omx.fast.scaled_dot_product_attention(qr, K[:, :, perm].contiguous(), V[:, :, perm].contiguous(), D ** -0.5)
    assert (o1.float() - o2.float()).abs().max().item() <= 1e-2
    assert (o1.float() - out.float()).abs().max().item() <= 1e-2
    o3 = omx.fast.scaled_dot_product_attention(qr, K, (V.float() * 2).bfloat16(), D ** -0.5)
    assert (o3.float() - 2 * o1.float()).abs().max().item() <= 2e-2


# ---- a composite call that fails leaves the cache exactly as it found it (offset, capacity, contents)
def _cache_fingerprint(gc):
    sk, sv = gc.state()
    return gc.offset(), tuple(sk.shape), sk.clone(), sv.clone()


def _same_cache(gc, fp):
    off, shape, k0, v0 = fp
    sk, sv = gc.state()
    return gc.offset() == off and tuple(sk.shape) == shape and torch.equal(sk, k0) and torch.equal(sv, v0)


@pytest.mark.parametrize("S", [255, 256])  # 256: the failing call would also have grown the cache
def test_failed_fused_decode_leaves_cache_unchanged(S):
    gc, oc = _prefill(2, 2, S, 128, "bf16", 31)
    fp = _cache_fingerprint(gc)
    q = randn((2, 8, 1, 128), "bf16", 1, DEV)
    k = randn((2, 2, 1, 128), "bf16", 2, DEV)
    v = randn((2, 2, 1, 128), "bf16", 3, DEV)
    rope = omx.nn.Rope(128, False, 1e6, 1.0)
    bad_out = torch.empty((2, 7, 1, 128), dtype=torch.bfloat16, device=DEV)  # wrong head count
    with pytest.raises(omx.Exception):
        omx.attn_decode_fused(q, k, v, gc, rope, 0.1, out=bad_out)
    assert _same_cache(gc, fp), "a failed fused decode advanced / grew the cache"
    q_bad = randn((2, 7, 1, 128), "bf16", 4, DEV)  # Hq % Hkv != 0
    with pytest.raises(omx.Exception):
        omx.attn_decode_fused(q_bad, k, v, gc, rope, 0.1)
    assert _same_cache(gc, fp)
    # ... and the next good step behaves as if nothing had happened
    _step(gc, oc, 2, 8, 2, 128, "bf16", ROPE, 40)


def test_failed_fused_prefill_leaves_cache_unchanged():
    gc, oc = _prefill(1, 2, 100, 128, "bf16", 33)
    fp = _cache_fingerprint(gc)
    q = randn((1, 8, 200, 128), "bf16", 1, DEV)
    k = randn((1, 2, 200, 128), "bf16", 2, DEV)
    v = randn((1, 2, 200, 128), "bf16", 3, DEV)
    rope = omx.nn.Rope(128, False, 1e6, 1.0)
    bad_out = torch.empty((1, 8, 199, 128), dtype=torch.bfloat16, device=DEV)
    with pytest.raises(omx.Exception):
        omx.attn_prefill_fused(q, k, v, gc, rope, 0.1, "causal", out=bad_out)
    assert _same_cache(gc, fp), "a failed fused prefill advanced / grew the cache"


def test_failed_first_update_forgets_the_latched_shape():
    gc = omx.KVCache()
    q = randn((1, 7, 1, 128), "bf16", 1, DEV)  # 7 q heads over 2 kv heads: rejected after the cache latched [1,2,*,128]
    k = randn((1, 2, 1, 128), "bf16", 2, DEV)
    with pytest.raises(omx.Exception):
        omx.attn_decode_fused(q, k, k, gc, None, 0.1)
    assert gc.offset() == 0
    k4 = randn((1, 4, 3, 64), "bf16", 3, DEV)  # a different geometry is still acceptable: nothing was latched
    gc.update_and_fetch(k4, k4)
    assert gc.offset() == 3 and tuple(gc.state()[0].shape) == (1, 4, 256, 64)


@pytest.mark.parametrize("B,Hq,Hkv,start", [(1, 8, 2, 190), (2, 32, 8, 1000)])
def test_back_to_back_steps_overlapped_launches(B, Hq, Hkv, start):
    """Consecutive fused steps on ONE cache with no host synchronisation in between: launch t+1 is dispatched
    while launch t still runs (programmatic dependent launch) and requests the cache rows older than launch t
    before its dependency wait -- row `offset - 1`, written by launch t, must still be read after it.  Every
    step's output and the final cache contents are compared with the oracle chain."""
    D, dtype, steps = 128, "bf16", 40
    gc, oc = _prefill(B, Hkv, start, D, dtype, 90)
    rope = omx.nn.Rope(*ROPE)
    scale = D ** -0.5
    qs = randn((steps, B, Hq, 1, D), dtype, 91)
    ks = randn((steps, B, Hkv, 1, D), dtype, 92)
    vs = randn((steps, B, Hkv, 1, D), dtype, 93)
    qd, kd, vd = qs.to(DEV), ks.to(DEV), vs.to(DEV)
    outs = torch.empty((steps, B, Hq, 1, D), dtype=torch.bfloat16, device=DEV)
    torch.cuda.synchronize()
    for i in range(steps):
        omx.attn_decode_fused(qd[i], kd[i], vd[i], gc, rope, scale, out=outs[i])
    torch.cuda.synchronize()
    assert omx.last_kernel() == "decode_hmma_tma"
    for i in range(steps):
        want = _oracle_step(oc, t2n(qs[i], dtype), t2n(ks[i], dtype), t2n(vs[i], dtype), dtype, ROPE, scale)
        assert_close(outs[i].float().cpu().numpy(), n2f(want, dtype), dtype, f"back-to-back step {i}")
    sk, sv = gc.state()
    assert_bits_equal(sk, oc.keys, dtype, "KV cache keys after back-to-back steps")
    assert_bits_equal(sv, oc.values, dtype, "KV cache values after back-to-back steps")
