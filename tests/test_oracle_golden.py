"""Pin the CPU oracle against every value-level vector the reference holds for
this path, and cross-check the unpinned parts independently."""
import numpy as np
import pytest
import torch

from conftest import load_oracle, n2f

orc = load_oracle()
from oracle import mlx_random  # noqa: E402


# --- reference golden vector: mlx-rs/src/nn/positional_encoding.rs:432-463 and
# --- mlx-rs/src/fast.rs:231-251 (seed 71, uniform(0,1,[2,8,16]), rope dims=8)

def _golden_input():
    return mlx_random.uniform_f32(mlx_random.RandomState(71), (2, 8, 16))


def test_reference_rng_restated_exactly():
    a = _golden_input()
    # the reference asserts +-2%; the restated key schedule reproduces them to f32 eps
    assert abs(float(a.mean(dtype=np.float32)) - 0.5082664489746094) < 2e-7
    assert abs(float(a.sum(dtype=np.float32)) - 130.1162109375) < 5e-5


def test_reference_rope_golden_vector():
    a = _golden_input()
    out = orc.rope(a, 8, False, 10000.0, 1.0, 0, dtype="f32")
    assert out.shape == (2, 8, 16) and out.dtype == np.float32
    mean, total = float(out.mean(dtype=np.float64)), float(out.sum(dtype=np.float64))
    assert abs(mean - 0.4562537670135498) < 5e-7      # reference tolerance: 9.1e-3
    assert abs(total - 116.80096435546875) < 1e-4     # reference tolerance: 2.3
    # the vector discriminates the pairing mode: traditional=True must NOT match
    alt = orc.rope(a, 8, True, 10000.0, 1.0, 0, dtype="f32")
    assert abs(float(alt.sum(dtype=np.float64)) - 116.80096435546875) > 0.1


# --- reference golden vector: mlx-rs/src/fast.rs:253-274 (seed 103, uniform(0,1,[2,8,16]), weight = ones, eps 1e-5)

def test_reference_rms_norm_golden_vector():
    a = mlx_random.uniform_f32(mlx_random.RandomState(103), (2, 8, 16))
    out = orc.rms_norm(a, np.ones(16, np.float32), 1e-5, dtype="f32")
    assert out.shape == (2, 8, 16) and out.dtype == np.float32
    assert abs(float(out.mean(dtype=np.float64)) - 0.87293875) < 5e-7   # reference tolerance: 1.7e-2
    assert abs(float(out.sum(dtype=np.float64)) - 223.47232) < 1e-4     # reference tolerance: 4.5
    x64 = a.astype(np.float64)
    np.testing.assert_allclose(out, x64 / np.sqrt((x64 ** 2).mean(-1, keepdims=True) + 1e-5), rtol=3e-7)


def test_rms_norm_weight_and_16bit_rounding_chain():
    rng = np.random.default_rng(3)
    x = rng.standard_normal((3, 5, 128)).astype(np.float32)
    w = rng.standard_normal(128).astype(np.float32)
    o = orc.rms_norm(x, w, 1e-6, dtype="f32")
    x64 = x.astype(np.float64)
    np.testing.assert_allclose(o, w * (x64 / np.sqrt((x64 ** 2).mean(-1, keepdims=True) + 1e-6)), rtol=2e-6, atol=1e-6)
    np.testing.assert_array_equal(orc.rms_norm(x, None, 1e-6, dtype="f32"),
                                  orc.rms_norm(x, np.ones(128, np.float32), 1e-6, dtype="f32"))
    # bf16: normalised value is rounded to bf16 BEFORE the weight multiply (astype, then multiply)
    xb, wb = orc.f32_to_bf16_bits(x), orc.f32_to_bf16_bits(w)
    ob = orc.bf16_bits_to_f32(orc.rms_norm(xb, wb, 1e-6, dtype="bf16"))
    xf, wf = orc.bf16_bits_to_f32(xb).astype(np.float64), orc.bf16_bits_to_f32(wb).astype(np.float64)
    y = orc.bf16_bits_to_f32(orc.f32_to_bf16_bits((xf / np.sqrt((xf ** 2).mean(-1, keepdims=True) + 1e-6)).astype(np.float32)))
    want = orc.bf16_bits_to_f32(orc.f32_to_bf16_bits((wf * y).astype(np.float32)))
    assert (ob != want).mean() < 0.01  # identical except where the f32 sum order flips a bf16 rounding


def test_rope_tail_copied_and_norm_preserved():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2, 3, 5, 16)).astype(np.float32)
    for trad in (False, True):
        o = orc.rope(x, 8, trad, 10000.0, 1.0, 7, dtype="f32")
        np.testing.assert_array_equal(o[..., 8:], x[..., 8:])
        np.testing.assert_allclose((o[..., :8] ** 2).sum(-1), (x[..., :8] ** 2).sum(-1), rtol=1e-5)


def test_rope_same_position_for_every_batch_row():
    # position = offset + t for all b (Metal quirk NOT reproduced)
    x = np.random.default_rng(1).standard_normal((1, 2, 1, 32)).astype(np.float32)
    xb = np.repeat(x, 5, axis=0)
    o = orc.rope(xb, 32, False, 1e6, 1.0, 1234, dtype="f32")
    for b in range(1, 5):
        np.testing.assert_array_equal(o[b], o[0])


def test_rope_freqs_equals_base_when_freqs_are_base_powers():
    x = np.random.default_rng(2).standard_normal((1, 2, 4, 16)).astype(np.float32)
    half = 8
    freqs = (10000.0 ** (np.arange(half) / half)).astype(np.float32)
    a = orc.rope(x, 16, False, None, 1.0, 3, freqs=freqs, dtype="f32")
    b = orc.rope(x, 16, False, 10000.0, 1.0, 3, dtype="f32")
    np.testing.assert_allclose(a, b, atol=2e-5)


def test_rope_ndim3_and_ndim5_shapes():
    x = np.random.default_rng(3).standard_normal((2, 6, 8)).astype(np.float32)
    o3 = orc.rope(x, 8, False, 10000.0, 1.0, 2, dtype="f32")
    o4 = orc.rope(x[:, None], 8, False, 10000.0, 1.0, 2, dtype="f32")
    np.testing.assert_array_equal(o3, o4[:, 0])
    x5 = np.random.default_rng(4).standard_normal((2, 2, 3, 6, 8)).astype(np.float32)
    o5 = orc.rope(x5, 8, False, 10000.0, 1.0, 2, dtype="f32")
    o4 = orc.rope(x5.reshape(2, 6, 6, 8), 8, False, 10000.0, 1.0, 2, dtype="f32")
    np.testing.assert_array_equal(o5.reshape(2, 6, 6, 8), o4)


def test_rope_bf16_rounds_after_every_op():
    rng = np.random.default_rng(5)
    xf = rng.standard_normal((1, 2, 3, 8)).astype(np.float32)
    xb = orc.f32_to_bf16_bits(xf)
    x = orc.bf16_bits_to_f32(xb)
    o = n2f(orc.rope(xb, 8, False, 10000.0, 1.0, 5, dtype="bf16"), "bf16")
    c, s = orc.rope_table(3, 8, 10000.0, 1.0, 5)
    r = lambda z: orc.bf16_bits_to_f32(orc.f32_to_bf16_bits(z))  # noqa: E731
    c, s = r(c), r(s)
    x1, x2 = x[..., :4], x[..., 4:]
    want = np.concatenate([r(r(x1 * c) - r(x2 * s)), r(r(x1 * s) + r(x2 * c))], -1)
    np.testing.assert_array_equal(o, want)


# --- sdpa: unpinned by the reference (fast.rs:301-331 is shape/dtype only) ---

@pytest.mark.parametrize("seq_len", [63, 129])
def test_sdpa_reference_shape_test(seq_len):
    # the reference's own test_fast_sdpa: B2 H24 Dk64, f32 and f16, shape + dtype
    B, H, Dk = 2, 24, 64
    rng = np.random.default_rng(seq_len)
    for dt, npdt in (("f32", np.float32), ("f16", np.float16)):
        q, k, v = (rng.standard_normal((B, H, seq_len, Dk)).astype(npdt) for _ in range(3))
        o = orc.sdpa(q, k, v, 1.0 / np.sqrt(Dk), None, dtype=dt)
        assert o.shape == (B, H, seq_len, Dk) and o.dtype == npdt


MASKS = ["none", "causal", "bool", "add"]


def _mk(B, Hq, Hkv, Lq, Lk, D, seed):
    rng = np.random.default_rng(seed)
    return (rng.standard_normal((B, Hq, Lq, D)).astype(np.float32),
            rng.standard_normal((B, Hkv, Lk, D)).astype(np.float32),
            rng.standard_normal((B, Hkv, Lk, D)).astype(np.float32))


def _mask(kind, Lq, Lk, rng, npdt=np.float32):
    if kind == "none":
        return None
    if kind == "causal":
        return "causal"
    if kind == "bool":
        m = rng.random((Lq, Lk)) > 0.3
        m[:, 0] = True
        return m
    return rng.standard_normal((1, 1, Lq, Lk)).astype(npdt)


@pytest.mark.parametrize("mask", MASKS)
@pytest.mark.parametrize("shape", [(2, 4, 2, 1, 37, 32), (1, 6, 6, 19, 19, 16), (1, 8, 2, 5, 40, 64)])
def test_sdpa_f32_vs_float64_twin_and_torch(mask, shape):
    B, Hq, Hkv, Lq, Lk, D = shape
    q, k, v = _mk(*shape, seed=11)
    m = _mask(mask, Lq, Lk, np.random.default_rng(12))
    scale = D ** -0.5
    got = orc.sdpa(q, k, v, scale, m, dtype="f32")
    want = orc.sdpa_numpy(q, k, v, scale, m)
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-5)
    # third opinion: torch CPU fp64
    G = Hq // Hkv
    tq, tk, tv = (torch.from_numpy(a).double() for a in (q, k, v))
    tk, tv = tk.repeat_interleave(G, 1), tv.repeat_interleave(G, 1)
    if mask == "causal":
        am = torch.from_numpy(orc.create_causal_mask(Lq, max(Lk - Lq, 0)))
    elif mask == "none":
        am = None
    elif mask == "bool":
        am = torch.from_numpy(m)
    else:
        am = torch.from_numpy(m).double()
    tw = torch.nn.functional.scaled_dot_product_attention(tq, tk, tv, attn_mask=am, scale=scale)
    np.testing.assert_allclose(got, tw.numpy(), rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("mask", MASKS)
def test_sdpa_bf16_within_stated_tolerance(mask):
    B, Hq, Hkv, Lq, Lk, D = 2, 8, 2, 7, 50, 64
    q, k, v = _mk(B, Hq, Hkv, Lq, Lk, D, seed=21)
    qb, kb, vb = (orc.f32_to_bf16_bits(a) for a in (q, k, v))
    m = _mask(mask, Lq, Lk, np.random.default_rng(22))
    mb = orc.f32_to_bf16_bits(m) if mask == "add" else m
    got = n2f(orc.sdpa(qb, kb, vb, D ** -0.5, mb, dtype="bf16"), "bf16")
    mf = orc.bf16_bits_to_f32(mb) if mask == "add" else m
    want = orc.sdpa_numpy(*(orc.bf16_bits_to_f32(a) for a in (qb, kb, vb)), D ** -0.5, mf)
    assert np.abs(got - want).max() <= 2e-2  # north_star: bf16 2e-2 max-abs


def test_sdpa_causal_is_bottom_right_aligned():
    B, H, Lq, Lk, D = 1, 2, 3, 9, 16
    q, k, v = _mk(B, H, H, Lq, Lk, D, seed=31)
    a = orc.sdpa(q, k, v, 0.25, "causal", dtype="f32")
    b = orc.sdpa(q, k, v, 0.25, orc.create_causal_mask(Lq, Lk - Lq), dtype="f32")
    np.testing.assert_array_equal(a, b)


def test_sdpa_fully_masked_row_uses_finfo_min():
    B, H, Lq, Lk, D = 1, 1, 2, 5, 8
    q, k, v = _mk(B, H, H, Lq, Lk, D, seed=41)
    m = np.ones((Lq, Lk), bool)
    m[1, :] = False
    o = orc.sdpa(q, k, v, 1.0, m, dtype="f32")
    np.testing.assert_allclose(o[0, 0, 1], v[0, 0].mean(0), rtol=1e-5, atol=1e-6)
    o2 = orc.sdpa(q, k, v, 1.0, m, dtype="f32", bool_fill_neg_inf=True)
    assert np.isnan(o2[0, 0, 1]).all()
    np.testing.assert_array_equal(o[0, 0, 0], o2[0, 0, 0])


def test_dit_manual_attention_equals_sdpa_without_mask_f32():
    # two softmaxes over shared [txt;img] K/V == one non-causal sdpa
    B, H, D, txt, img = 1, 3, 32, 5, 12
    q, k, v = _mk(B, H, H, txt + img, txt + img, D, seed=51)
    whole = orc.dit_attention(q, k, v, "f32", np.sqrt(D).astype(np.float32))
    img_o = orc.dit_attention(q[:, :, txt:], k, v, "f32", np.sqrt(D).astype(np.float32))
    txt_o = orc.dit_attention(q[:, :, :txt], k, v, "f32", np.sqrt(D).astype(np.float32))
    np.testing.assert_array_equal(whole[:, :, txt:], img_o)
    np.testing.assert_array_equal(whole[:, :, :txt], txt_o)
    s = orc.sdpa(q, k, v, D ** -0.5, None, dtype="f32")
    np.testing.assert_allclose(whole, s, rtol=1e-4, atol=1e-5)


def test_dit_rope_pairs():
    rng = np.random.default_rng(61)
    B, S, H, D = 2, 6, 3, 16
    x = rng.standard_normal((B, S, H, D)).astype(np.float32)
    ids = rng.integers(0, 20, (B, S, 4)).astype(np.float32)
    c, s = orc.klein_rope_freqs(ids, [4, 4, 4, 4], 2000.0)
    o = orc.dit_rope(x, c, s, "f32")
    x0, x1 = x[..., 0::2], x[..., 1::2]
    np.testing.assert_allclose(o[..., 0::2], x0 * c[:, :, None] - x1 * s[:, :, None], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(o[..., 1::2], x1 * c[:, :, None] + x0 * s[:, :, None], rtol=1e-6, atol=1e-6)


# --- KVCache: SURVEY.md Appendix A known-answer cases (cache.rs:134-194) ---

def _kv(n, seed, B=1, H=2, Dk=4, Dv=4):
    rng = np.random.default_rng(seed)
    return (rng.standard_normal((B, H, n, Dk)).astype(np.float32),
            rng.standard_normal((B, H, n, Dv)).astype(np.float32))


def _cap(c):
    return c.keys.shape[2]


def test_kvcache_appendix_a():
    c = orc.KVCache()
    k5, v5 = _kv(5, 0)
    kk, vv = c.update_and_fetch(k5, v5)
    assert (c.offset(), _cap(c)) == (5, 256)                      # A1
    np.testing.assert_array_equal(kk, k5)
    assert not c.keys[:, :, 5:].any()
    rows = [k5]
    for i in range(251):
        k1, v1 = _kv(1, 100 + i)
        rows.append(k1)
        kk, vv = c.update_and_fetch(k1, v1)
    assert (c.offset(), _cap(c)) == (256, 256)                    # A2
    np.testing.assert_array_equal(kk, np.concatenate(rows, 2))
    c.update_and_fetch(*_kv(1, 999))
    assert (c.offset(), _cap(c)) == (257, 512)                    # A3

    c = orc.KVCache()
    c.update_and_fetch(*_kv(300, 1))
    assert (c.offset(), _cap(c)) == (300, 512)                    # A4
    k2, v2 = _kv(300, 2)
    kk, vv = c.update_and_fetch(k2, v2)
    assert (c.offset(), _cap(c)) == (600, 812)                    # A5: not a multiple of 256
    np.testing.assert_array_equal(kk[:, :, 300:], k2)
    assert not c.keys[:, :, 600:].any()

    c = orc.KVCache()
    ka, va = _kv(300, 3)
    c.update_and_fetch(ka, va)
    c.reset()
    kb, vb = _kv(10, 4)
    kk, vv = c.update_and_fetch(kb, vb)
    assert (c.offset(), _cap(c)) == (10, 512)                     # A6
    assert kk.shape[2] == 10
    np.testing.assert_array_equal(c.keys[:, :, 10:300], ka[:, :, 10:300])  # stale rows kept

    c = orc.KVCache()
    c.update_and_fetch(*_kv(256, 5))
    c.reset()
    c.update_and_fetch(*_kv(300, 6))
    assert (c.offset(), _cap(c)) == (300, 768)                    # A7

    c = orc.KVCache()
    k, v = _kv(3, 7, Dk=8, Dv=4)
    kk, vv = c.update_and_fetch(k, v)
    assert c.keys.shape == (1, 2, 256, 8) and c.values.shape == (1, 2, 256, 4)  # A8


def test_kvcache_equals_naive_concat():
    c, n = orc.KVCache(), orc.ConcatKeyValueCache()
    rng = np.random.default_rng(8)
    for i, cnt in enumerate([7, 1, 1, 250, 300, 1, 513, 2]):
        k, v = _kv(cnt, 200 + i)
        a = c.update_and_fetch(k, v)
        b = n.update_and_fetch(k, v)
        np.testing.assert_array_equal(a[0], b[0])
        np.testing.assert_array_equal(a[1], b[1])
        assert c.offset() == n.offset()


def test_causal_mask_helpers():
    m = orc.create_causal_mask(3, 2)
    assert m.shape == (3, 5)
    assert m.tolist() == [[True, True, True, False, False],
                          [True, True, True, True, False],
                          [True, True, True, True, True]]
    w = orc.create_causal_mask(4, 0, window_size=1)
    assert w.tolist() == [[True, False, False, False], [True, True, False, False],
                          [False, True, True, False], [False, False, True, True]]
    assert orc.create_attention_mask(1) is None
    assert orc.create_attention_mask(5) == "causal"
    assert orc.create_attention_mask(5, cache_offset=3, return_array=True).shape == (5, 8)
