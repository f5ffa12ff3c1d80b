"""fast::rms_norm and the fused q_norm / k_norm decode prologue vs the oracle:
rms_norm outputs BIT-EXACT (same op order as the MLX CPU fallback), fused decode within tolerance
with a bit-exact KV cache."""
import numpy as np
import pytest
import torch

from conftest import assert_bits_equal, assert_close, load_oracle, load_pkg, n2f, randn, t2n

pytestmark = pytest.mark.gpu
omx = load_pkg()
orc = load_oracle()
DEV = "cuda"


@pytest.mark.parametrize("dtype", ["f32", "bf16", "f16"])
@pytest.mark.parametrize("shape", [(2, 8, 16), (3, 5, 128), (4, 7, 64), (2, 3, 80), (1, 2, 4096), (5, 1)])
def test_rms_norm_bit_exact(dtype, shape):
    x = randn(shape, dtype, 1)
    w = randn((shape[-1],), dtype, 2)
    got = omx.fast.rms_norm(x.to(DEV), w.to(DEV), 1e-5)
    assert omx.last_kernel() == "rms_norm"
    assert_bits_equal(got, orc.rms_norm(t2n(x, dtype), t2n(w, dtype), 1e-5, dtype=dtype), dtype, f"rms_norm {shape}")
    got = omx.fast.rms_norm(x.to(DEV), None, 1e-6)
    assert_bits_equal(got, orc.rms_norm(t2n(x, dtype), None, 1e-6, dtype=dtype), dtype, "rms_norm without weight")


def test_rms_norm_reference_golden_on_device():
    from oracle import mlx_random
    a = mlx_random.uniform_f32(mlx_random.RandomState(103), (2, 8, 16))  # mlx-rs/src/fast.rs:253-274
    out = omx.fast.rms_norm(torch.from_numpy(a).to(DEV), torch.ones(16, device=DEV), 1e-5)
    assert abs(float(out.double().mean()) - 0.87293875) < 5e-7
    assert abs(float(out.double().sum()) - 223.47232) < 1e-4


@pytest.mark.parametrize("dtype", ["bf16", "f32"])
def test_rms_norm_caller_layout_head_split_view(dtype):
    # q_proj output [B, L, H*D] -> reshape [B, L, H, D] -> transpose [B, H, L, D]  (model.rs:172-176)
    B, L, H, D = 2, 9, 4, 128
    x = randn((B, L, H, D), dtype, 3)
    w = randn((D,), dtype, 4)
    got = omx.nn.RmsNorm(w.to(DEV), 1e-6)(x.to(DEV).transpose(1, 2))
    assert got.shape == (B, H, L, D)
    want = orc.rms_norm(t2n(x.transpose(1, 2), dtype), t2n(w, dtype), 1e-6, dtype=dtype)
    assert_bits_equal(got, want, dtype, "rms_norm on a transposed view")


def test_rms_norm_errors():
    x = torch.zeros(2, 8, device=DEV)
    with pytest.raises(omx.Exception, match="same size as the last dimension"):
        omx.fast.rms_norm(x, torch.ones(7, device=DEV), 1e-5)
    with pytest.raises(omx.Exception, match="1 dimension"):
        omx.fast.rms_norm(x, torch.ones(2, 8, device=DEV), 1e-5)


def _oracle_step(oc, q, kn, vn, qw, kw, eps, dtype, rope, scale):
    """Attention::forward decode step with q_norm / k_norm, qwen3-mlx/src/model.rs:172-212."""
    qn = orc.rms_norm(q, qw, eps, dtype=dtype)
    kk = orc.rms_norm(kn, kw, eps, dtype=dtype)
    off = oc.offset()
    qr = orc.rope(qn, *rope, off, dtype=dtype)
    kr = orc.rope(kk, *rope, off, dtype=dtype)
    K, V = oc.update_and_fetch(kr, vn)
    return orc.sdpa(qr, np.ascontiguousarray(K), np.ascontiguousarray(V), scale, None, dtype=dtype)


@pytest.mark.parametrize("dtype,B,Hq,Hkv,D,S,kernel", [
    ("bf16", 3, 32, 8, 128, 1000, "decode_hmma_tma"),   # Qwen3-8B geometry
    ("bf16", 1, 16, 8, 128, 4500, "decode_hmma_tma"),   # Qwen3-0.6B geometry, several splits
    ("f32", 2, 16, 8, 128, 300, "decode_simt"),         # C1 dtype
    ("f16", 2, 4, 4, 256, 77, "decode_simt"),           # one query head per kv head, D = 256: CUDA-core split-K, fused
    ("f16", 2, 4, 4, 64, 77, "sdpa_mma"),               # narrow heads: prologue + mma.sync tiles
    ("f16", 2, 8, 2, 64, 77, "sdpa_mma"),               # grouped heads outside D = 128: unfused composition, mma.sync tiles
    ("bf16", 3, 16, 2, 256, 300, "sdpa_mma"),           # Qwen3.5 full-attention geometry (qwen3.5-35B-mlx/src/attention.rs)
])
def test_fused_decode_with_q_k_norm(dtype, B, Hq, Hkv, D, S, kernel):
    rope_t = (D if D < 256 else 64, False, 1e6, 1.0)  # head dim 256: Qwen3.5's partial rotary (first 64 features)
    eps = 1e-6
    k, v = randn((B, Hkv, S, D), dtype, 1), randn((B, Hkv, S, D), dtype, 2)
    q = randn((B, 1, Hq, D), dtype, 3).transpose(1, 2)       # caller layout: [B, L, H, D] viewed [B, H, L, D]
    kn = randn((B, 1, Hkv, D), dtype, 4).transpose(1, 2)
    vn = randn((B, 1, Hkv, D), dtype, 5).transpose(1, 2)
    qw, kw = 1 + 0.1 * randn((D,), "f32", 6), 1 + 0.1 * randn((D,), "f32", 7)
    from conftest import tdt
    qw, kw = qw.to(tdt(dtype)), kw.to(tdt(dtype))
    gc, oc = omx.KVCache(), orc.KVCache()
    gc.update_and_fetch(k.to(DEV), v.to(DEV))
    oc.update_and_fetch(t2n(k, dtype), t2n(v, dtype))
    rope = omx.nn.Rope(*rope_t)
    qn_m, kn_m = omx.nn.RmsNorm(qw.to(DEV), eps), omx.nn.RmsNorm(kw.to(DEV), eps)
    omx.launch_count(reset=True)
    got = omx.attn_decode_fused(q.to(DEV), kn.to(DEV), vn.to(DEV), gc, rope, D ** -0.5, q_norm=qn_m, k_norm=kn_m)
    torch.cuda.synchronize()
    assert omx.last_kernel() == kernel and (omx.launch_count() == 1 or kernel == "sdpa_mma")
    want = _oracle_step(oc, t2n(q, dtype), t2n(kn, dtype), t2n(vn, dtype), t2n(qw, dtype), t2n(kw, dtype), eps, dtype,
                        rope_t, D ** -0.5)
    assert_close(got.float().cpu().numpy(), n2f(want, dtype), dtype, "fused decode with q/k norm")
    sk, sv = gc.state()
    assert_bits_equal(sk, oc.keys, dtype, "KV keys after norm + rope + append")
    assert_bits_equal(sv, oc.values, dtype, "KV values")
    # the unfused spelling (5 library calls) lands on the same cache bits and the same output within tolerance
    gc2 = omx.KVCache()
    gc2.update_and_fetch(k.to(DEV), v.to(DEV))
    got2 = omx.attn_decode_unfused(q.to(DEV), kn.to(DEV), vn.to(DEV), gc2, rope, D ** -0.5, q_norm=qn_m, k_norm=kn_m)
    assert_bits_equal(gc2.state()[0], oc.keys, dtype, "KV keys, unfused chain")
    assert_close(got2.float().cpu().numpy(), n2f(want, dtype), dtype, "unfused decode with q/k norm")


def test_fused_norm_only_on_one_side_and_no_rope():
    dtype, B, Hq, Hkv, D, S = "bf16", 2, 8, 2, 128, 333
    k, v = randn((B, Hkv, S, D), dtype, 1), randn((B, Hkv, S, D), dtype, 2)
    q, kn, vn = randn((B, Hq, 1, D), dtype, 3), randn((B, Hkv, 1, D), dtype, 4), randn((B, Hkv, 1, D), dtype, 5)
    kw = randn((D,), dtype, 6)
    gc, oc = omx.KVCache(), orc.KVCache()
    gc.update_and_fetch(k.to(DEV), v.to(DEV))
    oc.update_and_fetch(t2n(k, dtype), t2n(v, dtype))
    got = omx.attn_decode_fused(q.to(DEV), kn.to(DEV), vn.to(DEV), gc, None, D ** -0.5,
                                k_norm=omx.nn.RmsNorm(kw.to(DEV), 1e-5))
    kk = orc.rms_norm(t2n(kn, dtype), t2n(kw, dtype), 1e-5, dtype=dtype)
    K, V = oc.update_and_fetch(kk, t2n(vn, dtype))
    want = orc.sdpa(t2n(q, dtype), np.ascontiguousarray(K), np.ascontiguousarray(V), D ** -0.5, None, dtype=dtype)
    assert_close(got.float().cpu().numpy(), n2f(want, dtype), dtype, "k_norm only, no rope")
    assert_bits_equal(gc.state()[0], oc.keys, dtype, "KV keys (k_norm only)")
