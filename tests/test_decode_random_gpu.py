"""Seeded random sweep of the fused decode step against the oracle's op-by-op chain: shapes that move the launch
between its variants (TMA / CUDA-core kernel, 1..many splits, cluster / L2 combine, 8- and 16-head groups, q/k norm
on or off, traditional / partial rope, one-wave / multi-wave grids).  Attention within tolerance, KV cache bit-exact."""
import numpy as np
import pytest
import torch

from conftest import assert_bits_equal, assert_close, load_oracle, load_pkg, n2f, randn, t2n

pytestmark = pytest.mark.gpu
omx = load_pkg()
orc = load_oracle()
DEV = "cuda"


def _case(seed):
    r = np.random.default_rng(1000 + seed)
    dtype = ["bf16", "f16", "f32"][int(r.integers(0, 3))]
    D = int(r.choice([64, 128, 128, 128]))
    Hkv = int(r.choice([1, 2, 8]))
    G = int(r.choice([1, 2, 4, 8, 16]))
    B = int(r.choice([1, 1, 2, 5]))
    S = int(r.choice([1, 40, 63, 64, 65, 300, 1000, 2500, 6000]))
    if dtype == "f32":
        S = min(S, 2500)  # keep the oracle fast
    norm = bool(r.integers(0, 2))
    rope = [None, (D, False, 1e6, 1.0), (D, True, 10000.0, 1.0), (D // 2, False, 10000.0, 0.5)][int(r.integers(0, 4))]
    steps = int(r.integers(1, 4))
    return dtype, D, Hkv, G, B, S, norm, rope, steps


@pytest.mark.parametrize("seed", range(28))
def test_random_fused_decode_vs_oracle(seed):
    dtype, D, Hkv, G, B, S, norm, rope_t, steps = _case(seed)
    Hq = Hkv * G
    k0, v0 = randn((B, Hkv, S, D), dtype, seed * 7 + 1), randn((B, Hkv, S, D), dtype, seed * 7 + 2)
    gc, oc = omx.KVCache(), orc.KVCache()
    gc.update_and_fetch(k0.to(DEV), v0.to(DEV))
    oc.update_and_fetch(t2n(k0, dtype), t2n(v0, dtype))
    qw, kw = randn((D,), dtype, seed * 7 + 3), randn((D,), dtype, seed * 7 + 4)
    qn = omx.nn.RmsNorm(qw.to(DEV), 1e-6) if norm else None
    kn = omx.nn.RmsNorm(kw.to(DEV), 1e-6) if norm else None
    rope = None if rope_t is None else omx.nn.Rope(*rope_t)
    scale = D ** -0.5
    for t in range(steps):
        q = randn((B, 1, Hq, D), dtype, 100 * seed + 3 * t).transpose(1, 2)
        k = randn((B, 1, Hkv, D), dtype, 100 * seed + 3 * t + 1).transpose(1, 2)
        v = randn((B, 1, Hkv, D), dtype, 100 * seed + 3 * t + 2).transpose(1, 2)
        got = omx.attn_decode_fused(q.to(DEV), k.to(DEV), v.to(DEV), gc, rope, scale, q_norm=qn, k_norm=kn)
        qo, ko = t2n(q, dtype), t2n(k, dtype)
        if norm:
            qo = orc.rms_norm(qo, t2n(qw, dtype), 1e-6, dtype=dtype)
            ko = orc.rms_norm(ko, t2n(kw, dtype), 1e-6, dtype=dtype)
        off = oc.offset()
        if rope_t is not None:
            qo = orc.rope(qo, rope_t[0], rope_t[1], rope_t[2], rope_t[3], off, dtype=dtype)
            ko = orc.rope(ko, rope_t[0], rope_t[1], rope_t[2], rope_t[3], off, dtype=dtype)
        K, V = oc.update_and_fetch(ko, t2n(v, dtype))
        want = orc.sdpa(qo, np.ascontiguousarray(K), np.ascontiguousarray(V), scale, None, dtype=dtype)
        what = f"seed {seed}: {dtype} B{B} Hq{Hq}/Hkv{Hkv} D{D} S{S}+{t} norm={norm} rope={rope_t} [{omx.last_kernel()}]"
        assert_close(got.float().cpu().numpy(), n2f(want, dtype), dtype, what)
        assert gc.offset() == oc.offset()
    sk, sv = gc.state()
    assert_bits_equal(sk, oc.keys, dtype, "KV keys")
    assert_bits_equal(sv, oc.values, dtype, "KV values")
