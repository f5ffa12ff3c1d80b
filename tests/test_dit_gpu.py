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
