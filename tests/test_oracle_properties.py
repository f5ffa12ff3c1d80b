"""Randomised (hypothesis) properties of the CPU oracle and of the host-side sharding logic -- no GPU.
The oracle is the checker of every GPU parity test, so its own invariants are worth more than a few fixed shapes:
agreement with the float64 twin for arbitrary shapes / masks / GQA ratios, rope as a rotation, the KVCache growth
contract under arbitrary call sequences, and the sequence-sharded log-sum-exp merge for arbitrary partitions."""
import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from conftest import load_oracle, load_pkg

orc = load_oracle()
omx = load_pkg()
par = omx.parallel

SET = settings(max_examples=40, deadline=None, derandomize=True)


@SET
@given(B=st.integers(1, 2), Hkv=st.integers(1, 3), G=st.sampled_from([1, 2, 4]), Lq=st.integers(1, 9),
       extra=st.integers(0, 40), D=st.sampled_from([8, 16, 64]), mask=st.sampled_from(["none", "causal", "bool", "add"]),
       dv_less=st.sampled_from([0, 0, 8]), seed=st.integers(0, 2 ** 16))
def test_sdpa_matches_float64_twin(B, Hkv, G, Lq, extra, D, mask, dv_less, seed):
    rng = np.random.default_rng(seed)
    Hq, Lk = Hkv * G, Lq + extra
    Dv = D - dv_less if D > dv_less else D  # values narrower than keys: the absorbed-MLA layout (glm-4.7-flash-mlx)
    q = rng.standard_normal((B, Hq, Lq, D)).astype(np.float32)
    k = rng.standard_normal((B, Hkv, Lk, D)).astype(np.float32)
    v = rng.standard_normal((B, Hkv, Lk, Dv)).astype(np.float32)
    if mask == "none":
        m = None
    elif mask == "causal":
        m = "causal"
    elif mask == "bool":
        m = rng.random((Lq, Lk)) > 0.4
        m[:, 0] = True  # no fully hidden row (that case has its own test)
    else:
        m = rng.standard_normal((B, 1, Lq, Lk)).astype(np.float32)
    got = orc.sdpa(q, k, v, D ** -0.5, m, dtype="f32")
    want = orc.sdpa_numpy(q, k, v, D ** -0.5, m)
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-5)
    # bf16 chain stays within the stated bar of the exact answer computed from the SAME bf16 inputs
    qb, kb, vb = (orc.f32_to_bf16_bits(a) for a in (q, k, v))
    mb = orc.f32_to_bf16_bits(m) if mask == "add" else m
    gb = orc.bf16_bits_to_f32(orc.sdpa(qb, kb, vb, D ** -0.5, mb, dtype="bf16"))
    wb = orc.sdpa_numpy(*(orc.bf16_bits_to_f32(a) for a in (qb, kb, vb)), D ** -0.5,
                        orc.bf16_bits_to_f32(mb) if mask == "add" else m)
    assert np.abs(gb - wb).max() <= 2e-2


@SET
@given(T=st.integers(1, 7), half=st.sampled_from([2, 4, 16]), tail=st.sampled_from([0, 8]), trad=st.booleans(),
       offset=st.integers(0, 5000), scale=st.sampled_from([1.0, 0.25]), seed=st.integers(0, 2 ** 16))
def test_rope_is_a_rotation_of_the_first_dims(T, half, tail, trad, offset, scale, seed):
    rng = np.random.default_rng(seed)
    dims, D = 2 * half, 2 * half + tail
    x = rng.standard_normal((2, 3, T, D)).astype(np.float32)
    o = orc.rope(x, dims, trad, 10000.0, scale, offset, dtype="f32")
    np.testing.assert_array_equal(o[..., dims:], x[..., dims:])  # tail copied
    pair = (lambda a: (a[..., 0:dims:2], a[..., 1:dims:2])) if trad else (lambda a: (a[..., :half], a[..., half:dims]))
    (x1, x2), (o1, o2) = pair(x), pair(o)
    np.testing.assert_allclose(o1 ** 2 + o2 ** 2, x1 ** 2 + x2 ** 2, rtol=2e-5, atol=1e-6)  # norm of every pair kept
    c, s = orc.rope_table(T, dims, 10000.0, scale, offset)
    np.testing.assert_allclose(o1, x1 * c - x2 * s, rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(o2, x1 * s + x2 * c, rtol=1e-6, atol=1e-6)
    # the same position for every batch row and head
    assert all(np.array_equal(orc.rope(x[b:b + 1, h:h + 1], dims, trad, 10000.0, scale, offset, dtype="f32")[0, 0],
                              o[b, h]) for b in range(2) for h in range(3))


@SET
@given(ops=st.lists(st.one_of(st.integers(1, 600), st.just("reset")), min_size=1, max_size=8),
       step=st.sampled_from([16, 256]), seed=st.integers(0, 2 ** 16))
def test_kvcache_growth_contract_under_any_call_sequence(ops, step, seed):
    """cache.rs:134-194 restated independently: capacity arithmetic, contents == naive concatenation since the
    last reset, never-written rows zero, stale rows kept after reset."""
    rng = np.random.default_rng(seed)
    c = orc.KVCache(step)
    off, cap, hist, ever = 0, 0, [], None
    for op in ops:
        if op == "reset":
            c.reset()
            off, hist = 0, []
            continue
        n = op
        k = rng.standard_normal((1, 2, n, 4)).astype(np.float32)
        v = rng.standard_normal((1, 2, n, 4)).astype(np.float32)
        if cap == 0 or off + n > cap:  # grow: trim to offset first unless offset is a multiple of step
            keep = cap if (cap == 0 or off % step == 0) else off
            cap = keep + -(-n // step) * step
        K, V = c.update_and_fetch(k, v)
        hist.append(k)
        off += n
        assert (c.offset(), c.keys.shape[2]) == (off, cap)
        np.testing.assert_array_equal(K, np.concatenate(hist, 2))
        assert K.shape[2] == off and V.shape[2] == off
        ever = max(ever or 0, off)
    if hist:
        assert not c.keys[:, :, max(ever, off):].any()  # rows never written are +0.0


@SET
@given(world=st.integers(1, 6), S=st.integers(1, 70), G=st.sampled_from([1, 4]), seed=st.integers(0, 2 ** 16))
def test_sequence_sharded_merge_equals_full_attention(world, S, G, seed):
    """Any partition of the keys by position % world, partial (normalised output, m, l) per rank, log-sum-exp merge
    == attention over all keys (parallel.merge_partials is the host restatement of omx_seqshard_merge)."""
    g = torch.Generator().manual_seed(seed)
    B, Hkv, D = 2, 2, 16
    Hq = Hkv * G
    q, k, v = (torch.randn(s, generator=g, dtype=torch.float64) for s in ((B, Hq, 1, D), (B, Hkv, S, D), (B, Hkv, S, D)))
    scale = D ** -0.5
    parts = []
    for r in range(world):
        rows = par.seq_shard_rows(S, world, r)
        if rows.numel() == 0:  # a rank without keys: (m, l) = (-inf, 0), output undefined
            p = torch.full((B, Hq, D + 2), float("nan"), dtype=torch.float64)
            p[..., -2], p[..., -1] = float("-inf"), 0.0
        else:
            kk, vv = k[:, :, rows].repeat_interleave(G, 1), v[:, :, rows].repeat_interleave(G, 1)
            s = torch.einsum("bhd,bhjd->bhj", q[:, :, 0], kk) * scale * 1.4426950408889634
            m = s.max(-1).values
            e = torch.exp2(s - m[..., None])
            l = e.sum(-1)
            p = torch.cat([torch.einsum("bhj,bhjd->bhd", e, vv) / l[..., None], m[..., None], l[..., None]], -1)
        parts.append(p)
    got = par.merge_partials(torch.stack(parts))
    want = torch.nn.functional.scaled_dot_product_attention(q, k.repeat_interleave(G, 1), v.repeat_interleave(G, 1),
                                                            scale=scale)[:, :, 0]
    torch.testing.assert_close(got, want, rtol=1e-9, atol=1e-10)


@SET
@given(batch=st.integers(0, 100), world=st.integers(1, 8))
def test_batch_shard_is_a_balanced_partition(batch, world):
    spans = [par.batch_shard(batch, world, r) for r in range(world)]
    assert sum(n for _, n in spans) == batch
    assert all(spans[i][0] + spans[i][1] == spans[i + 1][0] for i in range(world - 1))
    assert max(n for _, n in spans) - min(n for _, n in spans) <= 1
