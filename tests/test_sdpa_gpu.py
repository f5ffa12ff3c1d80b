"""fast::scaled_dot_product_attention on the B200 vs the CPU oracle (all mask modes, GQA,
strided views, dtypes) at the north_star tolerances: fp32 1e-4 relative, 16-bit 2e-2 max-abs."""
import numpy as np
import pytest
import torch

from conftest import assert_close, load_oracle, load_pkg, n2f, randn, t2n

pytestmark = pytest.mark.gpu
omx = load_pkg()
orc = load_oracle()
DEV = "cuda"
Causal = omx.fast.ScaledDotProductAttentionMask.Causal


def _mask(kind, B, Hq, Lq, Lk, dtype, seed):
    g = torch.Generator().manual_seed(seed)
    if kind == "none":
        return None, None
    if kind == "causal":
        return Causal, "causal"
    if kind == "bool2d":
        m = torch.rand((Lq, Lk), generator=g) > 0.3
        m[:, 0] = True
        return m.to(DEV), m.numpy()
    if kind == "bool4d":
        m = torch.rand((B, 1, Lq, Lk), generator=g) > 0.3
        m[..., 0] = True
        return m.to(DEV), m.numpy()
    if kind == "add":
        from conftest import tdt
        m = torch.randn((1, 1, Lq, Lk), generator=g).to(tdt(dtype))
        return m.to(DEV), t2n(m, dtype)
    raise ValueError(kind)


def _run(shape, dtype, mask_kind, seed=0, qview=None, kvview=None, force=None):
    B, Hq, Hkv, Lq, Lk, D = shape
    q = randn((B, Hq, Lq, D), dtype, seed + 1)
    k = randn((B, Hkv, Lk, D), dtype, seed + 2)
    v = randn((B, Hkv, Lk, D), dtype, seed + 3)
    gm, om = _mask(mask_kind, B, Hq, Lq, Lk, dtype, seed + 4)
    qd, kd, vd = q.to(DEV), k.to(DEV), v.to(DEV)
    if qview:
        qd = qview(qd)
    if kvview:
        kd, vd = kvview(kd), kvview(vd)
    scale = D ** -0.5
    if force:
        omx.force_kernel(force)
    try:
        got = omx.fast.scaled_dot_product_attention(qd, kd, vd, scale, gm)
    finally:
        omx.force_kernel("")
    assert got.shape == (B, Hq, Lq, D) and got.dtype == qd.dtype
    want = orc.sdpa(t2n(q, dtype), t2n(k, dtype), t2n(v, dtype), scale, om, dtype=dtype)
    assert_close(got.float().cpu().numpy(), n2f(want, dtype), dtype, f"sdpa {shape} {dtype} {mask_kind}")
    return got


MASKS = ["none", "causal", "bool2d", "bool4d", "add"]


@pytest.mark.parametrize("dtype", ["f32", "bf16", "f16"])
@pytest.mark.parametrize("mask", MASKS)
def test_generic_all_masks(dtype, mask):
    _run((2, 4, 2, 9, 37, 64), dtype, mask, force="sdpa_generic")


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
@pytest.mark.parametrize("shape", [(1, 6, 6, 19, 19, 16), (1, 8, 2, 5, 40, 128), (2, 2, 1, 33, 65, 80),
                                   (1, 2, 2, 3, 130, 256), (1, 3, 1, 7, 7, 24)])
def test_generic_shapes(dtype, shape):
    _run(shape, dtype, "causal", force="sdpa_generic")
    _run(shape, dtype, "none", force="sdpa_generic")


@pytest.mark.parametrize("seq_len", [63, 129, 400])
@pytest.mark.parametrize("dtype", ["f32", "f16"])
def test_reference_test_fast_sdpa_shapes(seq_len, dtype):
    # mlx-rs/src/fast.rs:301-331: B2 H24 Dk64, shape + dtype (here with values checked as well)
    _run((2, 24, 24, seq_len, seq_len, 64), dtype, "none")


def test_causal_bottom_right_alignment():
    shape = (1, 4, 2, 5, 29, 64)
    a = _run(shape, "f32", "causal")
    B, Hq, Hkv, Lq, Lk, D = shape
    q = randn((B, Hq, Lq, D), "f32", 1).to(DEV)
    k = randn((B, Hkv, Lk, D), "f32", 2).to(DEV)
    v = randn((B, Hkv, Lk, D), "f32", 3).to(DEV)
    m = omx.create_causal_mask(Lq, Lk - Lq, device=DEV)  # utils.rs:134-153
    b = omx.fast.scaled_dot_product_attention(q, k, v, D ** -0.5, m)
    np.testing.assert_allclose(a.cpu().numpy(), b.cpu().numpy(), rtol=1e-5, atol=1e-6)


def test_fully_masked_row_follows_finfo_min_rule():
    B, H, Lq, Lk, D = 1, 2, 3, 11, 32
    q, k, v = (randn((B, H, L, D), "f32", s).to(DEV) for L, s in ((Lq, 1), (Lk, 2), (Lk, 3)))
    m = torch.ones(Lq, Lk, dtype=torch.bool, device=DEV)
    m[1, :] = False
    o = omx.fast.scaled_dot_product_attention(q, k, v, 1.0, m)
    np.testing.assert_allclose(o[0, :, 1].cpu().numpy(), v[0].mean(1).cpu().numpy(), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_caller_layouts_strided(dtype):
    # q as transposed [B,L,H,D] storage; k/v as [:, :, :Lk] slices of a larger cache buffer
    B, Hq, Hkv, Lq, Lk, D = 2, 8, 2, 6, 50, 128
    q = randn((B, Lq, Hq, D), dtype, 1).to(DEV)
    kbuf = randn((B, Hkv, 256, D), dtype, 2).to(DEV)
    vbuf = randn((B, Hkv, 256, D), dtype, 3).to(DEV)
    qv, kv, vv = q.transpose(1, 2), kbuf[:, :, :Lk], vbuf[:, :, :Lk]
    got = omx.fast.scaled_dot_product_attention(qv, kv, vv, D ** -0.5, Causal)
    want = orc.sdpa(t2n(qv, dtype), t2n(kv, dtype), t2n(vv, dtype), D ** -0.5, "causal", dtype=dtype)
    assert_close(got.float().cpu().numpy(), n2f(want, dtype), dtype, "strided")
    # the caller's epilogue: transpose back + reshape (qwen3-mlx/src/model.rs:211-212) works on it
    assert got.transpose(1, 2).reshape(B, Lq, Hq * D).shape == (B, Lq, Hq * D)


def test_mask_arrays_variant_uses_first():
    q, k, v = (randn((1, 2, 4, 16), "f32", s).to(DEV) for s in (1, 2, 3))
    m = torch.rand(4, 4, device=DEV) > 0.5
    m[:, 0] = True
    a = omx.fast.scaled_dot_product_attention(q, k, v, 0.25, [m, ~m])
    b = omx.fast.scaled_dot_product_attention(q, k, v, 0.25, m)
    assert torch.equal(a, b)


def test_wrapper_of_mlx_rs_core_utils():
    q, k, v = (randn((1, 4, 6, 32), "f32", s).to(DEV) for s in (1, 2, 3))
    a = omx.scaled_dot_product_attention(q, k, v, None, 0.2, omx.SdpaMask.Causal)
    b = omx.fast.scaled_dot_product_attention(q, k, v, 0.2, Causal)
    assert torch.equal(a, b)


def test_validation_errors():
    f = omx.fast.scaled_dot_product_attention
    q = torch.zeros(1, 4, 3, 16, device=DEV)
    k = torch.zeros(1, 2, 5, 16, device=DEV)
    with pytest.raises(omx.Exception, match="matching last dimension"):
        f(q, torch.zeros(1, 2, 5, 8, device=DEV), torch.zeros(1, 2, 5, 8, device=DEV), 1.0)
    with pytest.raises(omx.Exception, match="multiple of n_kv_heads"):
        f(q, torch.zeros(1, 3, 5, 16, device=DEV), torch.zeros(1, 3, 5, 16, device=DEV), 1.0)
    with pytest.raises(omx.Exception, match="batch dimension"):
        f(q, torch.zeros(2, 2, 5, 16, device=DEV), torch.zeros(2, 2, 5, 16, device=DEV), 1.0)
    with pytest.raises(omx.Exception, match="not supported; expected"):
        f(q[0], k[0], k[0], 1.0)
    with pytest.raises(omx.Exception, match="unsupported type|Received unsupported"):
        f(q.int(), k.int(), k.int(), 1.0)
    with pytest.raises(omx.Exception, match="broadcastable"):
        f(q, k, k, 1.0, torch.ones(3, 4, dtype=torch.bool, device=DEV))
    with pytest.raises(omx.Exception, match="promote"):
        f(q.bfloat16(), k.bfloat16(), k.bfloat16(), 1.0, torch.zeros(3, 5, device=DEV))


@pytest.mark.parametrize("dtype", ["f32", "bf16", "f16"])
@pytest.mark.parametrize("shape", [(2, 8, 2, 1, 300, 128), (1, 16, 8, 1, 2048, 128), (3, 4, 4, 1, 77, 64),
                                   (1, 32, 2, 1, 1000, 128), (2, 6, 2, 1, 513, 128), (1, 8, 8, 1, 64, 256)])
def test_decode_dispatch_lq1(dtype, shape):
    # Lq == 1 goes to the split-K decode kernels (mask none; "causal" is the same thing at Lq == 1) -- except 16-bit
    # narrow heads (D <= 64), which the key-group mma.sync tiles serve faster
    want = "sdpa_mma" if dtype != "f32" and shape[5] <= 64 else "decode"
    _run(shape, dtype, "none")
    assert omx.last_kernel().startswith(want), omx.last_kernel()
    _run(shape, dtype, "causal")
    assert omx.last_kernel().startswith(want), omx.last_kernel()


def test_decode_kernel_families():
    _run((2, 8, 2, 1, 700, 128), "bf16", "none")
    assert omx.last_kernel() == "decode_hmma_tma"
    _run((2, 8, 2, 1, 700, 128), "f32", "none")
    assert omx.last_kernel() == "decode_simt"
    _run((2, 8, 2, 1, 700, 64), "bf16", "none")
    assert omx.last_kernel() == "sdpa_mma"          # grouped heads outside head dim 128: key-group mma.sync tiles
    _run((2, 4, 4, 1, 700, 64), "bf16", "none")
    assert omx.last_kernel() == "sdpa_mma"          # ... and narrow heads with one query head per kv head
    _run((2, 4, 4, 1, 700, 256), "bf16", "none")
    assert omx.last_kernel() == "decode_simt"       # wide heads, one query head per kv head: CUDA-core split-K kernel
    _run((2, 8, 2, 1, 700, 64), "bf16", "none", force="decode_simt")
    assert omx.last_kernel() == "decode_simt"
    _run((2, 8, 2, 1, 700, 128), "bf16", "bool2d")
    assert omx.last_kernel() == "decode_hmma_tma"   # array masks stay on the split-K decode kernels
    _run((2, 8, 2, 1, 700, 96), "bf16", "bool2d")
    assert omx.last_kernel() == "sdpa_mma"          # head dims outside {32, 64, 128, 256}: mma.sync tiles
    _run((2, 8, 2, 1, 700, 100), "bf16", "bool2d")
    assert omx.last_kernel() == "sdpa_generic"      # rows that are not 16-byte multiples: one warp per row


def test_decode_on_cache_views():
    # K/V fetched from KVCache are [..., :offset, :] views with head stride cap*D
    B, Hq, Hkv, S, D = 2, 8, 2, 333, 128
    c = omx.KVCache()
    k = randn((B, Hkv, S, D), "bf16", 1).to(DEV)
    v = randn((B, Hkv, S, D), "bf16", 2).to(DEV)
    kk, vv = c.update_and_fetch(k, v)
    assert kk.stride(1) == 512 * D
    q = randn((B, Hq, 1, D), "bf16", 3).to(DEV)
    got = omx.fast.scaled_dot_product_attention(q, kk, vv, D ** -0.5)
    assert omx.last_kernel() == "decode_hmma_tma"
    want = orc.sdpa(t2n(q, "bf16"), t2n(k, "bf16"), t2n(v, "bf16"), D ** -0.5, None, dtype="bf16")
    assert_close(got.float().cpu().numpy(), n2f(want, "bf16"), "bf16", "decode on cache views")


def test_empty_inputs():
    q = torch.zeros(0, 4, 3, 16, device=DEV)
    k = torch.zeros(0, 2, 5, 16, device=DEV)
    assert omx.fast.scaled_dot_product_attention(q, k, k, 1.0).shape == (0, 4, 3, 16)


# ---- decode (Lq == 1) with array masks stays on the split-K decode kernels (sliding windows,
# padding masks); rows whose mask hides every key follow the reference's finfo.min rule (uniform average)

def _decode_masked(B, Hq, Hkv, Lk, D, dtype, mask_t, expect):
    q = randn((B, Hq, 1, D), dtype, 1)
    k = randn((B, Hkv, Lk, D), dtype, 2)
    v = randn((B, Hkv, Lk, D), dtype, 3)
    scale = D ** -0.5
    got = omx.fast.scaled_dot_product_attention(q.to(DEV), k.to(DEV), v.to(DEV), scale, mask_t.to(DEV))
    torch.cuda.synchronize()
    assert omx.last_kernel() == expect, omx.last_kernel()
    om = mask_t.numpy() if mask_t.dtype == torch.bool else t2n(mask_t, dtype)
    want = orc.sdpa(t2n(q, dtype), t2n(k, dtype), t2n(v, dtype), scale, om, dtype=dtype)
    assert_close(got.float().cpu().numpy(), n2f(want, dtype), dtype, f"masked decode {dtype} D{D} Lk{Lk}")


@pytest.mark.parametrize("dtype,D,expect", [("bf16", 128, "decode_hmma_tma"), ("f16", 128, "decode_hmma_tma"),
                                            ("f32", 128, "decode_simt"), ("bf16", 64, "sdpa_mma"),
                                            ("f32", 64, "decode_simt")])
def test_decode_sliding_window_bool_mask(dtype, D, expect):
    Lk = 1500
    m = omx.create_causal_mask(1, Lk - 1, window_size=300, device="cpu")  # [1, Lk]: last 301 keys visible
    assert m.shape == (1, Lk) and int(m.sum()) == 301
    _decode_masked(2, 8, 2, Lk, D, dtype, m, expect)


@pytest.mark.parametrize("dtype,expect", [("bf16", "decode_hmma_tma"), ("f32", "decode_simt")])
def test_decode_per_batch_head_masks_and_hidden_rows(dtype, expect):
    g = torch.Generator().manual_seed(9)
    B, Hq, Hkv, Lk, D = 3, 8, 2, 777, 128
    m = torch.rand((B, Hq, 1, Lk), generator=g) > 0.6
    m[0, 3] = False          # one (batch, head) row sees nothing -> uniform average of V
    m[2] = False             # a whole batch item sees nothing
    m[1, :, :, :700] = False  # first tiles hidden entirely
    m[1, :, :, 701] = True
    _decode_masked(B, Hq, Hkv, Lk, D, dtype, m, expect)
    mb = torch.rand((B, 1, 1, Lk), generator=g) > 0.5   # padding-style [B,1,1,Lk]
    mb[..., 0] = True
    _decode_masked(B, Hq, Hkv, Lk, D, dtype, mb, expect)


@pytest.mark.parametrize("dtype,expect", [("bf16", "decode_hmma_tma"), ("f32", "decode_simt")])
def test_decode_additive_masks(dtype, expect):
    from conftest import tdt
    g = torch.Generator().manual_seed(10)
    B, Hq, Hkv, Lk, D = 2, 4, 4, 390, 128
    bias = (0.25 * torch.randn((1, Hq, 1, Lk), generator=g)).to(tdt(dtype))
    _decode_masked(B, Hq, Hkv, Lk, D, dtype, bias, expect)
    keep = torch.rand((B, 1, 1, Lk), generator=g) > 0.5
    keep[1] = False   # padded-out batch item
    add = ((~keep).float() * -1e9).to(tdt(dtype))
    _decode_masked(B, Hq, Hkv, Lk, D, dtype, add, expect)


def test_decode_mask_long_context_split_k():
    # single sequence, long context: the split-K plan with the last-CTA combine, masked
    Lk = 20000
    m = omx.create_causal_mask(1, Lk - 1, window_size=4095, device="cpu")
    _decode_masked(1, 32, 8, Lk, 128, "bf16", m, "decode_hmma_tma")
    none = torch.zeros((1, Lk), dtype=torch.bool)
    _decode_masked(1, 32, 8, Lk, 128, "bf16", none, "decode_hmma_tma")


@pytest.mark.parametrize("L", [1, 7])
def test_absorbed_mla_glm47_flash_dk576_dv512(L):
    # GLM-4.7-Flash absorbed MLA (glm-4.7-flash-mlx/src/model.rs:263-299): ONE shared kv head, keys
    # [B,1,S,512+64], values [B,1,S,512] (a view of the keys' first 512 features in the crate), 20 query heads
    B, H, S, Dk, Dv, dtype = 1, 20, 300, 576, 512, "bf16"
    q = randn((B, H, L, Dk), dtype, 1)
    k = randn((B, 1, S, Dk), dtype, 2)
    v = k[..., :Dv]  # strided view, like the crate's `values = kv_latent`
    mask = None if L == 1 else omx.fast.ScaledDotProductAttentionMask.Causal
    got = omx.fast.scaled_dot_product_attention(q.to(DEV), k.to(DEV), k.to(DEV)[..., :Dv], Dk ** -0.5, mask)
    assert tuple(got.shape) == (B, H, L, Dv) and omx.last_kernel() == "sdpa_mma"
    want = orc.sdpa(t2n(q, dtype), t2n(k, dtype), t2n(v.contiguous(), dtype), Dk ** -0.5,
                    None if L == 1 else "causal", dtype=dtype)
    assert_close(got.float().cpu().numpy(), n2f(want, dtype), dtype, "absorbed MLA attention")
    # through the cache, Dk != Dv
    c = omx.KVCache()
    K, V = c.update_and_fetch(k.to(DEV), k.to(DEV)[..., :Dv].contiguous())
    got2 = omx.fast.scaled_dot_product_attention(q.to(DEV), K, V, Dk ** -0.5, mask)
    assert torch.equal(got, got2)


# ---- 16-bit shapes outside the tcgen05 / TMA kernels: mma.sync tiles with the query group packed (sdpa_mma.cu) ----

def _run_wide(B, Hq, Hkv, Lq, Lk, Dk, Dv, dtype, mask_kind, seed=0, force="sdpa_mma", v_is_k_view=False):
    q = randn((B, Hq, Lq, Dk), dtype, seed + 1)
    k = randn((B, Hkv, Lk, Dk), dtype, seed + 2)
    v = k[..., :Dv] if v_is_k_view else randn((B, Hkv, Lk, Dv), dtype, seed + 3)
    gm, om = _mask(mask_kind, B, Hq, Lq, Lk, dtype, seed + 4)
    kd = k.to(DEV)
    vd = kd[..., :Dv] if v_is_k_view else v.to(DEV)
    scale = Dk ** -0.5
    omx.force_kernel(force)
    try:
        got = omx.fast.scaled_dot_product_attention(q.to(DEV), kd, vd, scale, gm)
        assert omx.last_kernel() == force
    finally:
        omx.force_kernel("")
    assert tuple(got.shape) == (B, Hq, Lq, Dv)
    want = orc.sdpa(t2n(q, dtype), t2n(k, dtype), t2n(v.contiguous(), dtype), scale, om, dtype=dtype)
    assert_close(got.float().cpu().numpy(), n2f(want, dtype), dtype,
                 f"sdpa_mma {(B, Hq, Hkv, Lq, Lk, Dk, Dv)} {dtype} {mask_kind}")
    return got


@pytest.mark.parametrize("dtype", ["bf16", "f16"])
@pytest.mark.parametrize("mask", MASKS)
def test_mma_all_masks(dtype, mask):
    """Forced kernel vs the oracle chain: ragged rows / keys, GQA groups packed into the row tile (group size that
    does not divide 64), Lq < Lk (bottom-right causal), every mask mode; head dims 80 (padded to 96) and 256."""
    _run_wide(2, 6, 2, 37, 150, 80, 80, dtype, mask)
    _run_wide(1, 4, 2, 70, 70, 256, 256, dtype, mask)


@pytest.mark.parametrize("dims", [(16, 16), (24, 24), (72, 72), (96, 96), (128, 128), (192, 128), (256, 256),
                                  (576, 512), (320, 264)])
def test_mma_head_dims(dims):
    Dk, Dv = dims
    _run_wide(1, 6, 3, 33, 97, Dk, Dv, "bf16", "causal", seed=Dk)
    _run_wide(2, 4, 1, 5, 200, Dk, Dv, "f16", "none", seed=Dk + 1)


@pytest.mark.parametrize("L", [1, 3, 64, 130])
def test_mma_absorbed_mla_decode_and_prefill(L):
    """GLM-4.7-Flash absorbed MLA (glm-4.7-flash-mlx/src/model.rs:263-299): 20 heads over ONE latent kv head,
    keys 512 + 64, values = the keys' first 512 features (a strided view); L = 1 splits the keys over CTAs and
    merges in a second launch."""
    S = 700 + L
    _run_wide(2, 20, 1, L, S, 576, 512, "bf16", "none" if L == 1 else "causal", v_is_k_view=True)
    got = _run_wide(1, 20, 1, L, S, 576, 512, "bf16", "bool2d")
    ref = _run_wide(1, 20, 1, L, S, 576, 512, "bf16", "bool2d", force="sdpa_generic")
    np.testing.assert_allclose(got.float().cpu().numpy(), ref.float().cpu().numpy(), atol=1.5e-2)


@pytest.mark.parametrize("H,L,S", [(8, 1, 333), (16, 1, 64), (20, 1, 31), (32, 1, 1000), (16, 2, 500), (33, 1, 257)])
def test_mma_mla_decode_key_groups(H, L, S):
    """Packed rows <= 32 (decode) run the variant whose warps split each key tile into two key groups merged at the
    end; 33 rows and up take the four-row-group layout.  Ragged key counts, bool mask, several batches."""
    _run_wide(3, H, 1, L, S, 576, 512, "bf16", "none", seed=H + S)
    _run_wide(2, H, 1, L, S, 576, 512, "f16", "bool4d", seed=H + S + 1)


@pytest.mark.parametrize("D", [32, 64, 80, 128, 256])
@pytest.mark.parametrize("Hq,Hkv,L", [(4, 4, 1), (8, 2, 1), (16, 2, 1), (32, 2, 1), (8, 2, 2), (8, 1, 5)])
def test_mma_decode_key_groups_every_width(D, Hq, Hkv, L):
    """Few packed rows (G x L <= 16: four key groups per tile, <= 32: two) on every width configuration: ragged
    key counts incl. fewer keys than one tile, key range split over CTAs at the longer context, causal for L > 1."""
    mask = "none" if L == 1 else "causal"
    _run_wide(2, Hq, Hkv, L, 45, D, D, "bf16", mask, seed=D + Hq)
    _run_wide(3, Hq, Hkv, L, 1500 + 7 * L, D, D, "f16", "bool4d" if L == 1 else mask, seed=D + Hq + 1)


def test_mma_seeded_random_sweep_vs_generic_and_oracle():
    """60 seeded random calls over the kernel's whole domain -- widths 8 .. 576 / 512 (values never wider than keys),
    group sizes 1 .. 20, 1 .. 70 query rows, ragged key counts, every mask mode, both 16-bit types -- forced onto
    sdpa_mma and compared with the row-per-warp kernel (same inputs, float32 arithmetic) and, for the small ones,
    with the oracle chain."""
    rng = np.random.default_rng(20261018)
    widths = [8, 16, 24, 40, 64, 72, 80, 96, 104, 128, 192, 256, 320, 576]
    for it in range(60):
        B, Hkv = int(rng.integers(1, 4)), int(rng.integers(1, 4))
        G = int(rng.choice([1, 2, 3, 4, 8, 16, 20]))
        Lq = int(rng.choice([1, 1, 2, 3, 7, 17, 33, 70]))
        Lk = Lq + int(rng.integers(0, 300))
        Dk = int(rng.choice(widths))
        Dv = min(512, int(rng.choice([Dk, Dk, max(8, Dk - 8), max(8, (Dk // 16) * 8)])))
        dtype = "bf16" if it % 2 else "f16"
        mask = MASKS[int(rng.integers(0, len(MASKS)))]
        Hq = Hkv * G
        q = randn((B, Hq, Lq, Dk), dtype, 1000 + it)
        k = randn((B, Hkv, Lk, Dk), dtype, 2000 + it)
        v = randn((B, Hkv, Lk, Dv), dtype, 3000 + it)
        gm, om = _mask(mask, B, Hq, Lq, Lk, dtype, 4000 + it)
        outs = {}
        for kern in ("sdpa_mma", "sdpa_generic"):
            omx.force_kernel(kern)
            try:
                outs[kern] = omx.fast.scaled_dot_product_attention(q.to(DEV), k.to(DEV), v.to(DEV), Dk ** -0.5, gm)
                assert omx.last_kernel() == kern
            finally:
                omx.force_kernel("")
        what = f"#{it} {(B, Hq, Hkv, Lq, Lk, Dk, Dv)} {dtype} {mask}"
        a, b = outs["sdpa_mma"].float().cpu().numpy(), outs["sdpa_generic"].float().cpu().numpy()
        assert np.isfinite(a).all(), what
        assert np.abs(a - b).max() <= 2e-2, f"{what}: kernels differ by {np.abs(a - b).max():.3e}"
        if B * Hq * Lq * Lk * Dk <= 4e7:
            want = orc.sdpa(t2n(q, dtype), t2n(k, dtype), t2n(v, dtype), Dk ** -0.5, om, dtype=dtype)
            assert_close(a, n2f(want, dtype), dtype, what)


def test_mma_large_batch_grid():
    # the batch index rides grid.z, (row tile, key split) grid.x: batches beyond a few hundred stay on this kernel
    _run_wide(700, 2, 1, 1, 40, 80, 80, "bf16", "none")
    _run_wide(300, 4, 2, 3, 70, 576, 512, "f16", "causal")


def test_mma_split_keys_long_context_and_masked_rows():
    # one packed row tile, 9000 keys: split over the SMs; bool rows that hide every key follow the finfo.min rule
    _run_wide(1, 20, 1, 1, 9000, 576, 512, "bf16", "none")
    _run_wide(1, 8, 2, 1, 5000, 80, 80, "f16", "add")
    B, H, Lq, Lk, D = 1, 4, 40, 300, 80
    q, k, v = (randn((B, H, L, D), "bf16", s) for L, s in ((Lq, 1), (Lk, 2), (Lk, 3)))
    m = torch.rand((Lq, Lk), generator=torch.Generator().manual_seed(5)) > 0.5
    m[7, :] = False
    omx.force_kernel("sdpa_mma")
    try:
        got = omx.fast.scaled_dot_product_attention(q.to(DEV), k.to(DEV), v.to(DEV), D ** -0.5, m.to(DEV))
    finally:
        omx.force_kernel("")
    want = orc.sdpa(t2n(q, "bf16"), t2n(k, "bf16"), t2n(v, "bf16"), D ** -0.5, m.numpy(), dtype="bf16")
    assert_close(got.float().cpu().numpy(), n2f(want, "bf16"), "bf16", "fully masked row")
    np.testing.assert_allclose(got[0, :, 7].float().cpu().numpy(), v[0].float().mean(1).numpy(), atol=1e-2)


def test_mma_strided_views_and_refusals():
    # q stored [B, L, H, D]; k / v as slices of a longer cache buffer (the KVCache views)
    B, Hq, Hkv, Lq, Lk, D, dtype = 2, 8, 2, 9, 77, 80, "bf16"
    q = randn((B, Lq, Hq, D), dtype, 1)
    kb, vb = randn((B, Hkv, 256, D), dtype, 2), randn((B, Hkv, 256, D), dtype, 3)
    got = omx.fast.scaled_dot_product_attention(q.to(DEV).transpose(1, 2), kb.to(DEV)[:, :, :Lk], vb.to(DEV)[:, :, :Lk],
                                                D ** -0.5, Causal)
    assert omx.last_kernel() == "sdpa_mma"
    want = orc.sdpa(t2n(q.transpose(1, 2), dtype), t2n(kb[:, :, :Lk], dtype), t2n(vb[:, :, :Lk], dtype), D ** -0.5,
                    "causal", dtype=dtype)
    assert_close(got.float().cpu().numpy(), n2f(want, dtype), dtype, "strided sdpa_mma")
    omx.force_kernel("sdpa_mma")
    try:
        with pytest.raises(omx.Exception, match="not bf16 / f16"):
            x = torch.zeros((1, 2, 32, 80), dtype=torch.float32, device=DEV)
            omx.fast.scaled_dot_product_attention(x, x, x, 1.0, None)
        with pytest.raises(omx.Exception, match="multiples of 8"):
            x = torch.zeros((1, 2, 32, 100), dtype=torch.bfloat16, device=DEV)
            omx.fast.scaled_dot_product_attention(x, x, x, 1.0, None)
    finally:
        omx.force_kernel("")


# ---- float32 with many query rows: the tiled FFMA kernel (sdpa_f32_tiled.cu) ----

@pytest.mark.parametrize("mask", MASKS)
@pytest.mark.parametrize("shape", [(2, 4, 2, 70, 70, 128), (1, 4, 4, 64, 200, 64), (1, 6, 2, 130, 257, 128)])
def test_f32_tiled_all_masks(shape, mask):
    """Forced kernel vs the oracle chain at the float32 bar (1e-4 relative): ragged row / key counts, GQA, the
    bottom-right aligned causal mask with Lq < Lk, bool and additive arrays, both head dims."""
    _run(shape, "f32", mask, force="sdpa_f32_tiled")


def test_f32_tiled_is_the_default_for_float32_prefill_and_matches_generic():
    shape = (1, 16, 8, 300, 300, 128)  # Qwen3-0.6B-shape prefill (C1's model), float32
    got = _run(shape, "f32", "causal")
    assert omx.last_kernel() == "sdpa_f32_tiled"
    ref = _run(shape, "f32", "causal", force="sdpa_generic")
    assert omx.last_kernel() == "sdpa_generic"
    np.testing.assert_allclose(got.cpu().numpy(), ref.cpu().numpy(), rtol=2e-5, atol=2e-6)
    # a handful of rows stays on the row-per-warp kernel; 16-bit inputs never come here
    _run((1, 4, 2, 8, 100, 128), "f32", "causal")
    assert omx.last_kernel() == "sdpa_generic"
    omx.force_kernel("sdpa_f32_tiled")
    try:
        with pytest.raises(omx.Exception, match="not float32"):
            q = torch.zeros((1, 2, 32, 128), dtype=torch.bfloat16, device=DEV)
            omx.fast.scaled_dot_product_attention(q, q, q, 1.0, None)
    finally:
        omx.force_kernel("")


def test_f32_tiled_strided_views_and_fully_masked_rows():
    # [B, L, H, D] projections viewed [B, H, L, D]; K / V as a slice of a longer cache buffer
    _run((2, 4, 2, 90, 90, 128), "f32", "causal", force="sdpa_f32_tiled",
         qview=lambda t: t.transpose(1, 2).contiguous().transpose(1, 2),
         kvview=lambda t: torch.cat([t, torch.zeros_like(t)], 2)[:, :, :t.shape[2]])
    # rows whose bool mask hides every key follow the finfo.min rule (uniform average), like the generic kernel
    B, H, L, D = 1, 2, 80, 64
    q, k, v = (randn((B, H, L, D), "f32", s) for s in (1, 2, 3))
    m = torch.rand((L, L), generator=torch.Generator().manual_seed(5)) > 0.5
    m[7, :] = False
    m[70, :] = False
    omx.force_kernel("sdpa_f32_tiled")
    try:
        got = omx.fast.scaled_dot_product_attention(q.to(DEV), k.to(DEV), v.to(DEV), D ** -0.5, m.to(DEV))
    finally:
        omx.force_kernel("")
    want = orc.sdpa(t2n(q, "f32"), t2n(k, "f32"), t2n(v, "f32"), D ** -0.5, m.numpy(), dtype="f32")
    assert_close(got.cpu().numpy(), n2f(want, "f32"), "f32", "fully masked rows, tiled kernel")
