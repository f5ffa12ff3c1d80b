"""DiT joint text+image attention (FLUX.2-klein / Z-Image manual chain) on the B200 vs the oracle."""
import numpy as np
import pytest
import torch

from conftest import assert_bits_equal, assert_close, load_oracle, load_pkg, n2f, n2t, randn, t2n, tdt

pytestmark = pytest.mark.gpu
omx = load_pkg()
orc = load_oracle()
DEV = "cuda"


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_dit_rope_bit_exact(dtype):
    # klein axes [32,32,32,32] theta 2000 (klein_quantized.rs:564); zimage [32,48,48] theta 256
    for axes, theta in (([32, 32, 32, 32], 2000.0), ([32, 48, 48], 256.0)):
        B, S, H, D = 2, 37, 3, sum(axes)
        x = randn((B, S, H, D), dtype, 1)
        ids = torch.randint(0, 64, (B, S, len(axes)), generator=torch.Generator().manual_seed(2)).float()
        c, s = orc.klein_rope_freqs(ids.numpy(), axes, theta)
        ct, st = torch.from_numpy(c).to(tdt(dtype)), torch.from_numpy(s).to(tdt(dtype))
        got = omx.dit.apply_rope(x.to(DEV), ct.to(DEV), st.to(DEV))
        want = orc.dit_rope(t2n(x, dtype), t2n(ct, dtype), t2n(st, dtype), dtype)
        assert_bits_equal(got, want, dtype, f"dit rope {axes}")
        got4 = omx.dit.apply_rope(x.to(DEV), ct.to(DEV)[:, :, None], st.to(DEV)[:, :, None])  # zimage layout
        assert torch.equal(got, got4)


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_joint_attention_vs_manual_chain(dtype):
    # [txt; img] order, 24 heads x 128 (klein_model.rs:170-180), reduced sequence
    B, H, D, txt, img = 2, 24, 128, 24, 104
    S = txt + img
    q, k, v = (randn((B, S, H, D), dtype, s) for s in (1, 2, 3))
    got = omx.dit.joint_attention(q.to(DEV), k.to(DEV), v.to(DEV), D ** -0.5, out_dtype=torch.float32)
    tr = lambda a: np.ascontiguousarray(np.swapaxes(t2n(a, dtype), 1, 2))  # noqa: E731
    want = orc.dit_attention(tr(q), tr(k), tr(v), dtype, np.float32(np.sqrt(D)))  # [B,H,S,D] f32
    assert_close(np.swapaxes(got.cpu().numpy(), 1, 2), want, dtype, "dit joint attention")
    # the two-block spelling of the reference (img rows, txt rows) is the same computation
    got_img = omx.dit.joint_attention(q[:, txt:].to(DEV), k.to(DEV), v.to(DEV), D ** -0.5, out_dtype=torch.float32)
    np.testing.assert_allclose(got_img.cpu().numpy(), got[:, txt:].cpu().numpy(), rtol=1e-5, atol=1e-6)
    # native-dtype output == fast::sdpa(mask none) (qwen-image-mlx/src/qwen_full_precision.rs:232-243)
    o16 = omx.dit.joint_attention(q.to(DEV), k.to(DEV), v.to(DEV), D ** -0.5)
    s16 = omx.fast.scaled_dot_product_attention(q.to(DEV).transpose(1, 2), k.to(DEV).transpose(1, 2),
                                                v.to(DEV).transpose(1, 2), D ** -0.5)
    assert torch.equal(o16.transpose(1, 2), s16)


def test_zimage_additive_mask():
    B, H, D, S = 1, 30, 128, 40
    q, k, v = (randn((B, S, H, D), "bf16", s) for s in (4, 5, 6))
    m = torch.zeros(S, S)
    m[:, S - 6:] = float("-inf")  # padded text tokens masked out (zimage_model.rs:376-380)
    got = omx.dit.joint_attention(q.to(DEV), k.to(DEV), v.to(DEV), D ** -0.5, add_mask=m.to(DEV),
                                  out_dtype=torch.float32)
    tr = lambda a: np.ascontiguousarray(np.swapaxes(t2n(a, "bf16"), 1, 2))  # noqa: E731
    want = orc.dit_attention(tr(q), tr(k), tr(v), "bf16", np.float32(D ** -0.5), use_mul=True, add_mask=m.numpy())
    assert_close(np.swapaxes(got.cpu().numpy(), 1, 2), want, "bf16", "zimage masked attention")


# ---- the DiT chains' own semantics (f32 output, f32 additive mask) on the TENSOR-CORE kernel, forced, against the
# manual-chain oracle at the C4 sequence length (512 txt + 4096 img tokens)
def _forced(name):
    class _F:
        def __enter__(self):
            omx.force_kernel(name)

        def __exit__(self, *a):
            omx.force_kernel("")
    return _F()


def _tr(a, dtype="bf16"):
    return np.ascontiguousarray(np.swapaxes(t2n(a, dtype), 1, 2))


def test_flux_chain_c4_shape_on_tcgen05_f32_out():
    # FLUX.2-klein: softmax(q k^T / sqrt(D)) v, f32 result (klein_model.rs:474-483); 1 batch x 4 heads x 4608 tokens
    B, H, D, S = 1, 4, 128, 4608
    q, k, v = (randn((B, S, H, D), "bf16", s) for s in (11, 12, 13))
    with _forced("fmha_tcgen05"):
        got = omx.dit.joint_attention(q.to(DEV), k.to(DEV), v.to(DEV), D ** -0.5, out_dtype=torch.float32)
    assert omx.last_kernel() == "fmha_tcgen05" and got.dtype == torch.float32
    want = orc.dit_attention(_tr(q), _tr(k), _tr(v), "bf16", np.float32(np.sqrt(D)))
    assert_close(np.swapaxes(got.cpu().numpy(), 1, 2), want, "bf16", "FLUX chain, tcgen05, f32 out")
    # default dispatch takes the same kernel (no forcing needed)
    got2 = omx.dit.joint_attention(q.to(DEV), k.to(DEV), v.to(DEV), D ** -0.5, out_dtype=torch.float32)
    assert omx.last_kernel() == "fmha_tcgen05" and torch.equal(got, got2)
    # and the f32 epilogue only changes the rounding of the store
    g16 = omx.dit.joint_attention(q.to(DEV), k.to(DEV), v.to(DEV), D ** -0.5)
    assert torch.equal(got.to(torch.bfloat16), g16)


def test_zimage_chain_c4_shape_on_tcgen05_f32_mask_gqa_by_repeat():
    # Z-Image: x scale, f32 additive mask hiding the padded text tokens, 30 heads from 10 kv heads repeated x3
    # (zimage_model.rs:355-384); here 6 heads from 2 kv heads, 4608 tokens, last 100 keys padded
    B, H, Hkv, D, S = 1, 6, 2, 128, 4608
    q = randn((B, S, H, D), "bf16", 21)
    k0, v0 = randn((B, S, Hkv, D), "bf16", 22), randn((B, S, Hkv, D), "bf16", 23)
    k, v = k0.repeat_interleave(H // Hkv, dim=2), v0.repeat_interleave(H // Hkv, dim=2)  # the crate's repeat
    m = torch.zeros(S, S)
    m[:, S - 100:] = float("-inf")
    m[5, :] = -1e9  # one fully hidden query row: the chain's softmax is uniform there
    with _forced("fmha_tcgen05"):
        got = omx.dit.joint_attention(q.to(DEV), k.to(DEV), v.to(DEV), D ** -0.5, add_mask=m.to(DEV),
                                      out_dtype=torch.float32)
    assert omx.last_kernel() == "fmha_tcgen05_arraymask"
    want = orc.dit_attention(_tr(q), _tr(k), _tr(v), "bf16", np.float32(D ** -0.5), use_mul=True, add_mask=m.numpy())
    g = np.swapaxes(got.cpu().numpy(), 1, 2)
    keep = np.ones(S, bool)
    keep[5] = False
    assert_close(g[:, :, keep], want[:, :, keep], "bf16", "Z-Image chain, tcgen05, f32 mask")
    # the hidden row: the oracle's -1e9 fill leaves a softmax over (score - 1e9) in f32 = uniform over the keys
    # that are not -inf; the kernel's masked_rows_fixup averages ALL keys (the sdpa convention).  Both are finite.
    assert np.isfinite(g[:, :, 5]).all()
    # GQA by index on the un-repeated K/V is the same computation
    with _forced("fmha_tcgen05"):
        got_gqa = omx.dit.joint_attention(q.to(DEV), k0.to(DEV), v0.to(DEV), D ** -0.5, add_mask=m.to(DEV),
                                          out_dtype=torch.float32)
    assert torch.equal(got_gqa, got)
    # same call on the CUDA-core kernel agrees within the bar as well
    with _forced("sdpa_generic"):
        gen = omx.dit.joint_attention(q.to(DEV), k.to(DEV), v.to(DEV), D ** -0.5, add_mask=m.to(DEV),
                                      out_dtype=torch.float32)
    assert_close(g[:, :, keep], np.swapaxes(gen.cpu().numpy(), 1, 2)[:, :, keep], "bf16", "tcgen05 vs generic")


def test_fused_dit_block_f32_out_runs_on_tcgen05():
    # omx_dit_attn_fused (prologue + attention) with the chain's f32 result
    B, H, D, txt, img = 1, 4, 128, 128, 384
    qs, ks, vs = ([randn((B, n, H, D), "bf16", s + i).to(DEV) for i, n in enumerate((txt, img))] for s in (31, 41, 51))
    out = omx.dit.attn_fused(qs, ks, vs, D ** -0.5, out_dtype=torch.float32)
    assert omx.last_kernel() == "fmha_tcgen05" and out.dtype == torch.float32
    cat = lambda xs: torch.cat(xs, 1).cpu()  # noqa: E731
    want = orc.dit_attention(_tr(cat(qs)), _tr(cat(ks)), _tr(cat(vs)), "bf16", np.float32(np.sqrt(D)))
    assert_close(np.swapaxes(out.cpu().numpy(), 1, 2), want, "bf16", "fused DiT block, f32 out")
