"""tcgen05/TMEM/TMA flash-attention forward (causal prefill + non-causal DiT) vs the CPU oracle."""
import numpy as np
import pytest
import torch

from conftest import assert_close, load_oracle, load_pkg, n2f, randn, t2n

pytestmark = pytest.mark.gpu
omx = load_pkg()
orc = load_oracle()
DEV = "cuda"
Causal = omx.fast.ScaledDotProductAttentionMask.Causal


def _run(B, Hq, Hkv, Lq, Lk, dtype="bf16", causal=True, seed=0, qview=False, kvslice=False, tol_check=True, D=128):
    if qview:  # [B, L, H, D] storage viewed [B, H, L, D]
        q = randn((B, Lq, Hq, D), dtype, seed + 1).transpose(1, 2)
    else:
        q = randn((B, Hq, Lq, D), dtype, seed + 1)
    cap = Lk + 77 if kvslice else Lk
    kbuf = randn((B, Hkv, cap, D), dtype, seed + 2)
    vbuf = randn((B, Hkv, cap, D), dtype, seed + 3)
    qd, kd, vd = q.to(DEV), kbuf.to(DEV)[:, :, :Lk], vbuf.to(DEV)[:, :, :Lk]
    scale = D ** -0.5
    omx.force_kernel("fmha_tcgen05")
    try:
        got = omx.fast.scaled_dot_product_attention(qd, kd, vd, scale, Causal if causal else None)
    finally:
        omx.force_kernel("")
    torch.cuda.synchronize()
    assert omx.last_kernel() == "fmha_tcgen05"
    assert got.shape == (B, Hq, Lq, D)
    want = orc.sdpa(t2n(q, dtype), t2n(kbuf[:, :, :Lk], dtype), t2n(vbuf[:, :, :Lk], dtype), scale,
                    "causal" if causal else None, dtype=dtype)
    if tol_check:
        assert_close(got.float().cpu().numpy(), n2f(want, dtype), dtype,
                     f"fmha B{B} Hq{Hq} Hkv{Hkv} Lq{Lq} Lk{Lk} causal={causal}")
    return got, want


@pytest.mark.parametrize("causal", [False, True])
def test_single_tile(causal):
    _run(1, 1, 1, 128, 128, causal=causal)


@pytest.mark.parametrize("causal", [False, True])
def test_two_q_tiles_two_kv_tiles(causal):
    _run(1, 2, 2, 256, 256, causal=causal)


@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("L", [384, 1024])
def test_square_gqa(causal, L):
    _run(2, 8, 2, L, L, causal=causal)


@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("Lq,Lk", [(100, 100), (129, 257), (300, 300), (511, 777), (64, 1000), (257, 130)])
def test_ragged_lengths(causal, Lq, Lk):
    # partial Q / KV tiles; causal is bottom-right aligned when Lq < Lk, top-left when Lq > Lk
    _run(1, 4, 2, Lq, Lk, causal=causal)


def test_f16():
    _run(1, 4, 4, 256, 384, dtype="f16", causal=True)
    _run(1, 4, 4, 256, 384, dtype="f16", causal=False)


def test_caller_layouts():
    _run(2, 8, 2, 300, 300, causal=True, qview=True, kvslice=True)


def test_long_kv_many_ring_wraps():
    _run(1, 2, 1, 256, 4096, causal=False)
    _run(1, 2, 1, 512, 4096, causal=True)


def test_c4_dit_shape_reduced_batch():
    # FLUX.2-klein: 24 heads x 128, 512 txt + 4096 img tokens (C4), one batch item, 4 heads checked
    got, want = _run(1, 4, 4, 4608, 4608, causal=False)


def test_c3_prefill_shape_slice():
    # Qwen3-8B prefill geometry (GQA 4:1, seq 8192) on one kv head group
    _run(1, 4, 1, 8192, 8192, causal=True)


def test_auto_dispatch_and_bool_mask_equivalence():
    B, Hq, Hkv, L, D = 1, 8, 2, 512, 128
    q, k, v = (randn((B, h, L, D), "bf16", s).to(DEV) for h, s in ((Hq, 1), (Hkv, 2), (Hkv, 3)))
    a = omx.fast.scaled_dot_product_attention(q, k, v, D ** -0.5, Causal)
    assert omx.last_kernel() == "fmha_tcgen05"
    m = omx.create_causal_mask(L, 0, device=DEV)  # the array mask the LLM crates build (utils.rs:134-153)
    b = omx.fast.scaled_dot_product_attention(q, k, v, D ** -0.5, m)
    assert omx.last_kernel() == "fmha_tcgen05_arraymask"  # array masks stay on the tensor-core path
    assert (a.float() - b.float()).abs().max().item() <= 1e-2


# ---- array masks on the tcgen05 path (what the LLM crates' prefill actually passes: qwen3-mlx/src/model.rs:401,
# mixtral-mlx/src/model.rs:388 call create_attention_mask(h, cache, Some(true)) -> a bool [T, offset+T] array)

def _run_arr(B, Hq, Hkv, Lq, Lk, mask_t, dtype="bf16", seed=0, expect="fmha_tcgen05_arraymask", D=128):
    q = randn((B, Hq, Lq, D), dtype, seed + 1)
    k = randn((B, Hkv, Lk, D), dtype, seed + 2)
    v = randn((B, Hkv, Lk, D), dtype, seed + 3)
    scale = D ** -0.5
    got = omx.fast.scaled_dot_product_attention(q.to(DEV), k.to(DEV), v.to(DEV), scale, mask_t.to(DEV))
    torch.cuda.synchronize()
    assert omx.last_kernel() == expect, omx.last_kernel()
    om = mask_t.numpy() if mask_t.dtype == torch.bool else t2n(mask_t, dtype)
    want = orc.sdpa(t2n(q, dtype), t2n(k, dtype), t2n(v, dtype), scale, om, dtype=dtype)
    assert_close(got.float().cpu().numpy(), n2f(want, dtype), dtype, f"fmha array mask Lq{Lq} Lk{Lk}")
    return got


@pytest.mark.parametrize("T,offset", [(128, 0), (300, 0), (512, 256), (777, 100), (1024, 3000)])
def test_bool_causal_array_mask_like_create_causal_mask(T, offset):
    m = omx.create_causal_mask(T, offset, device="cpu")  # [T, offset+T]
    _run_arr(1, 4, 2, T, offset + T, m)


def test_bool_window_mask_skips_tiles_on_both_sides():
    m = omx.create_causal_mask(1024, 512, window_size=200, device="cpu")
    _run_arr(1, 2, 1, 1024, 1536, m)


def test_bool_mask_broadcast_shapes_and_random_pattern():
    g = torch.Generator().manual_seed(5)
    B, Hq, Lq, Lk = 2, 4, 300, 520
    m4 = torch.rand((B, 1, Lq, Lk), generator=g) > 0.3
    m4[..., 0] = True  # every row keeps at least one key
    _run_arr(B, Hq, 2, Lq, Lk, m4)
    mh = torch.rand((1, Hq, Lq, Lk), generator=g) > 0.5
    mh[..., 5] = True
    _run_arr(B, Hq, 2, Lq, Lk, mh)
    m2 = torch.rand((Lq, Lk), generator=g) > 0.9
    m2[:, 100] = True
    _run_arr(B, Hq, 2, Lq, Lk, m2)


@pytest.mark.parametrize("dtype", ["bf16", "f16"])
def test_additive_array_masks(dtype):
    from conftest import tdt
    g = torch.Generator().manual_seed(6)
    Lq, Lk = 260, 390
    # dense additive bias; kept small because the reference rounds (scores + mask) to the 16-bit dtype before
    # the softmax, and with O(1) biases that rounding noise alone reaches the 2e-2 bar at Lk ~ 400
    m = (0.25 * torch.randn((1, 1, Lq, Lk), generator=g)).to(tdt(dtype))
    _run_arr(1, 4, 4, Lq, Lk, m, dtype=dtype)
    # (1 - mask) * -1e9 in the q dtype, the way zimage-mlx builds it (qwen3_quantized.rs:148-164)
    keep = omx.create_causal_mask(384, 0, device="cpu")
    add = ((~keep).float() * -1e9).to(tdt(dtype))
    _run_arr(1, 2, 2, 384, 384, add, dtype=dtype)
    neg_inf = torch.where(keep, 0.0, float("-inf")).to(tdt(dtype))
    _run_arr(1, 2, 2, 384, 384, neg_inf, dtype=dtype)


def test_array_mask_unaligned_rows_and_fully_masked_tiles():
    # Lk not a multiple of 16: mask rows are not 16-byte aligned -> byte path; plus a row block whose
    # first KV tiles are entirely masked
    g = torch.Generator().manual_seed(7)
    Lq, Lk = 200, 333
    m = torch.rand((Lq, Lk), generator=g) > 0.4
    m[:, :128] = False
    m[:, 200] = True
    _run_arr(1, 2, 2, Lq, Lk, m)


def test_rows_without_visible_keys_average_v_like_the_reference():
    # degenerate rows (no visible key): the reference's CPU chain fills with finfo.min, so the softmax is
    # uniform over ALL keys; the tile-skipping kernel flags such rows and masked_rows_fixup rewrites them
    Lq = Lk = 256
    m = omx.create_causal_mask(Lq, 0, device="cpu").clone()
    m[10] = False
    m[200:] = False   # a whole 128-row tile's worth of rows and more
    _run_arr(1, 2, 2, Lq, Lk, m)
    m4 = m[None, None].repeat(2, 4, 1, 1).clone()
    m4[1, 2] = True   # per-(batch, head) pattern: one head sees everything
    m4[0, 0] = False  # one head sees nothing: every KV tile skipped for its CTAs
    _run_arr(2, 4, 2, Lq, Lk, m4)


@pytest.mark.parametrize("dtype", ["bf16", "f16"])
def test_additive_mask_rows_without_visible_keys(dtype):
    # left-padded batch the way flux-klein's text encoder masks it (qwen3_encoder.rs:173-200):
    # causal AND key-not-padding, as (1 - keep) * -1e9 in the q dtype; padded query rows see no key at all
    from conftest import tdt
    L, pad = 300, 37
    keep = omx.create_causal_mask(L, 0, device="cpu").clone()
    keep[:, :pad] = False
    add = ((~keep).float() * -1e9).to(tdt(dtype))[None, None]
    if dtype == "bf16":
        _run_arr(1, 4, 2, L, L, add, dtype=dtype)
        return
    # f16 cannot hold -1e9: the mask is -inf and the reference's softmax of an all -inf row is NaN.  This
    # library returns the uniform average there too (finite); every other row must match.
    q, k, v = (randn((1, h, L, 128), dtype, s) for h, s in ((4, 1), (2, 2), (2, 3)))
    got = omx.fast.scaled_dot_product_attention(q.to(DEV), k.to(DEV), v.to(DEV), 128 ** -0.5, add.to(DEV))
    assert omx.last_kernel() == "fmha_tcgen05_arraymask"
    want = n2f(orc.sdpa(t2n(q, dtype), t2n(k, dtype), t2n(v, dtype), 128 ** -0.5, t2n(add, dtype), dtype=dtype), dtype)
    assert np.isnan(want[:, :, :pad]).all() and torch.isfinite(got.float()).all()
    assert_close(got.float().cpu().numpy()[:, :, pad:], want[:, :, pad:], dtype, "rows with visible keys")


def test_c3_shape_bool_array_mask_matches_causal_string():
    # Qwen3-8B prefill geometry with the ARRAY mask the crate builds; equal work to "causal" thanks to tile skipping
    B, Hq, Hkv, L, D = 1, 8, 2, 4096, 128
    q, k, v = (randn((B, h, L, D), "bf16", s).to(DEV) for h, s in ((Hq, 1), (Hkv, 2), (Hkv, 3)))
    a = omx.fast.scaled_dot_product_attention(q, k, v, D ** -0.5, Causal)
    m = omx.create_causal_mask(L, 0, device=DEV)
    b = omx.fast.scaled_dot_product_attention(q, k, v, D ** -0.5, m)
    assert omx.last_kernel() == "fmha_tcgen05_arraymask"
    assert torch.equal(a, b)  # same tiles, same arithmetic order -> identical bits


def test_full_size_properties_c3():
    # BASELINE C3 at full size (B8, 32q/8kv, seq 8192, bf16): size-independent properties
    B, Hq, Hkv, L, D = 8, 32, 8, 8192, 128
    g = torch.Generator(device=DEV).manual_seed(1237)
    q = torch.randn((B, Hq, L, D), generator=g, device=DEV).bfloat16()
    k = torch.randn((B, Hkv, L, D), generator=g, device=DEV).bfloat16()
    v = torch.randn((B, Hkv, L, D), generator=g, device=DEV).bfloat16()
    o = omx.fast.scaled_dot_product_attention(q, k, v, D ** -0.5, Causal)
    assert omx.last_kernel() == "fmha_tcgen05"
    assert torch.isfinite(o.float()).all()
    # (1) row 0 attends to key 0 only: O[:, h, 0] == V[:, h/4, 0]
    assert torch.equal(o[:, :, 0], v[:, :, 0].repeat_interleave(Hq // Hkv, 1))
    # (2) causality: perturbing the future does not change the past
    k2, v2 = k.clone(), v.clone()
    k2[:, :, 4096:] = -k2[:, :, 4096:]
    v2[:, :, 4096:] = 0
    o2 = omx.fast.scaled_dot_product_attention(q, k2, v2, D ** -0.5, Causal)
    assert torch.equal(o[:, :, :4096], o2[:, :, :4096])
    # (3) linearity in V
    o3 = omx.fast.scaled_dot_product_attention(q, k, (v.float() * 2).bfloat16(), D ** -0.5, Causal)
    assert (o3.float() - 2 * o.float()).abs().max().item() <= 2e-2
    # (4) one (batch, kv-group) slice against the oracle
    want = orc.sdpa(t2n(q[3:4, 8:12], "bf16"), t2n(k[3:4, 2:3], "bf16"), t2n(v[3:4, 2:3], "bf16"), D ** -0.5,
                    "causal", dtype="bf16")
    assert_close(o[3:4, 8:12].float().cpu().numpy(), n2f(want, "bf16"), "bf16", "C3 slice")


def _oracle_rows(q, k, v, triples, G, causal, q_off=0):
    """Oracle output of sampled (batch, q head, query row) triples of a full-size launch: one Lq = 1 call each
    against the keys that row sees (bottom-right aligned causal mask = a prefix of the keys)."""
    D = q.shape[-1]
    out = []
    for b, h, r in triples:
        n = (q_off + r + 1) if causal else k.shape[2]
        want = orc.sdpa(t2n(q[b:b + 1, h:h + 1, r:r + 1], "bf16"), t2n(k[b:b + 1, h // G:h // G + 1, :n], "bf16"),
                        t2n(v[b:b + 1, h // G:h // G + 1, :n], "bf16"), D ** -0.5, None, dtype="bf16")
        out.append(n2f(want, "bf16")[0, 0, 0])
    return np.stack(out)


def test_full_size_c3_sampled_rows_against_the_oracle():
    # the FULL C3 launch (B8, 32q/8kv, seq 8192, causal), 64 random (batch, head, row) triples against the oracle:
    # every persistent CTA / work item / diagonal tile position is a candidate, not one fixed slice
    B, Hq, Hkv, L, D = 8, 32, 8, 8192, 128
    g = torch.Generator(device=DEV).manual_seed(4237)
    q = torch.randn((B, Hq, L, D), generator=g, device=DEV).bfloat16()
    k = torch.randn((B, Hkv, L, D), generator=g, device=DEV).bfloat16()
    v = torch.randn((B, Hkv, L, D), generator=g, device=DEV).bfloat16()
    o = omx.fast.scaled_dot_product_attention(q, k, v, D ** -0.5, Causal)
    assert omx.last_kernel() == "fmha_tcgen05"
    rng = np.random.default_rng(11)
    triples = [(int(rng.integers(B)), int(rng.integers(Hq)), int(rng.integers(L))) for _ in range(56)]
    triples += [(0, 0, 0), (7, 31, L - 1), (3, 5, 127), (3, 5, 128), (4, 17, 255), (4, 17, 256), (6, 30, 4095), (6, 30, 4096)]
    got = np.stack([o[b, h, r].float().cpu().numpy() for b, h, r in triples])
    assert_close(got, _oracle_rows(q, k, v, triples, Hq // Hkv, True), "bf16", "C3 full size, sampled rows")


def test_full_size_c4_sampled_rows_against_the_oracle():
    # the FULL C4 launch (B4, 24 heads, 512 txt + 4096 img), 64 random (batch, head, row) triples
    B, H, S, D = 4, 24, 4608, 128
    g = torch.Generator(device=DEV).manual_seed(4238)
    q, k, v = (torch.randn((B, S, H, D), generator=g, device=DEV).bfloat16().transpose(1, 2) for _ in range(3))
    o = omx.fast.scaled_dot_product_attention(q, k, v, D ** -0.5, None)
    assert omx.last_kernel() == "fmha_tcgen05"
    rng = np.random.default_rng(12)
    triples = [(int(rng.integers(B)), int(rng.integers(H)), int(rng.integers(S))) for _ in range(60)]
    triples += [(0, 0, 0), (3, 23, S - 1), (1, 7, 511), (1, 7, 512)]
    got = np.stack([o[b, h, r].float().cpu().numpy() for b, h, r in triples])
    assert_close(got, _oracle_rows(q, k, v, triples, 1, False), "bf16", "C4 full size, sampled rows")


def test_full_size_properties_c4():
    # BASELINE C4 at full size (FLUX.2-klein joint attention: B4, 24 heads, 512 txt + 4096 img tokens, bf16)
    B, H, S, D = 4, 24, 4608, 128
    g = torch.Generator(device=DEV).manual_seed(1238)
    # [B, S, H, D] storage viewed [B, H, S, D], the crates' layout (klein_model.rs:465-468)
    q, k, v = (torch.randn((B, S, H, D), generator=g, device=DEV).bfloat16().transpose(1, 2) for _ in range(3))
    o = omx.fast.scaled_dot_product_attention(q, k, v, D ** -0.5, None)
    assert omx.last_kernel() == "fmha_tcgen05" and torch.isfinite(o.float()).all()
    # (1) non-causal attention does not care about the order of the keys: permute [txt; img] rows of K and V together
    perm = torch.randperm(S, generator=torch.Generator().manual_seed(3)).to(DEV)
    o_p = omx.fast.scaled_dot_product_attention(q, k[:, :, perm], v[:, :, perm], D ** -0.5, None)
    assert (o.float() - o_p.float()).abs().max().item() <= 2e-2
    # (2) a query block's result does not depend on which other query rows ride in the launch
    o_txt = omx.fast.scaled_dot_product_attention(q[:, :, :512], k, v, D ** -0.5, None)
    assert torch.equal(o_txt, o[:, :, :512])
    # (3) convexity: every output lies inside the per-feature range of V
    vmin, vmax = v.float().amin(2, keepdim=True), v.float().amax(2, keepdim=True)
    assert bool(((o.float() >= vmin - 1e-2) & (o.float() <= vmax + 1e-2)).all())
    # (4) two heads of one batch item against the oracle
    want = orc.sdpa(t2n(q[2:3, 5:7], "bf16"), t2n(k[2:3, 5:7], "bf16"), t2n(v[2:3, 5:7], "bf16"), D ** -0.5, None,
                    dtype="bf16")
    assert_close(o[2:3, 5:7].float().cpu().numpy(), n2f(want, "bf16"), "bf16", "C4 slice")


# ---- head_dim 64 on the same kernel (Qwen3-ASR audio encoder: embed_dim / heads = 64, non-causal,
# qwen3-asr-mlx/src/encoder.rs:133-180; and the shapes of the reference's own sdpa tests, mlx-rs/src/fast.rs:301-331)

@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("Lq,Lk", [(128, 128), (256, 384), (63, 63), (129, 129), (400, 400), (300, 1000)])
def test_head_dim_64(causal, Lq, Lk):
    _run(2, 4, 2, Lq, Lk, causal=causal, D=64)


def test_head_dim_64_reference_test_shape_f16():
    # fast.rs:301-331: B2, H24, D64, f16
    for L in (63, 129, 400):
        _run(2, 24, 24, L, L, dtype="f16", causal=False, D=64, seed=L)


def test_head_dim_64_caller_layouts_and_long_kv():
    _run(2, 8, 2, 300, 300, causal=True, qview=True, kvslice=True, D=64)
    _run(1, 2, 1, 512, 4096, causal=True, D=64)
    _run(1, 20, 20, 1500, 1500, causal=False, D=64)   # audio-encoder-sized window


def test_head_dim_64_array_masks():
    m = omx.create_causal_mask(512, 256, window_size=200, device="cpu")
    _run_arr(1, 4, 2, 512, 768, m, D=64)
    m2 = omx.create_causal_mask(300, 0, device="cpu").clone()
    m2[17] = False
    _run_arr(1, 2, 2, 300, 300, m2, D=64)
