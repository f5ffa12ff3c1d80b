"""CUDA-graph decode loop: the fused decode step with the position read on the device must be
BIT-identical to the host-offset step (same kernels, same arithmetic) -- outputs and the whole KV cache --
whether launched eagerly or replayed from a captured graph, and stay within tolerance of the oracle."""
import numpy as np
import pytest
import torch

from conftest import assert_bits_equal, assert_close, load_oracle, load_pkg, n2f, randn, t2n

pytestmark = pytest.mark.gpu
omx = load_pkg()
orc = load_oracle()
DEV = "cuda"


def _caches(B, Hkv, S, Dk, dtype, seed, n=2):
    k, v = randn((B, Hkv, S, Dk), dtype, seed), randn((B, Hkv, S, Dk), dtype, seed + 1)
    out = []
    for _ in range(n):
        c = omx.KVCache()
        c.update_and_fetch(k.to(DEV), v.to(DEV))
        out.append(c)
    return out, k, v


CASES = [  # B, Hq, Hkv, S0, D, dtype, steps, norm, kernel
    (1, 16, 8, 300, 128, "f32", 5, False, "decode_simt"),         # C1 shape, split-K on CUDA cores
    (1, 32, 8, 1000, 128, "bf16", 6, True, "decode_hmma_tma"),    # single sequence: many splits + combine
    (4, 8, 2, 250, 128, "bf16", 10, False, "decode_hmma_tma"),    # crosses the 256-row growth boundary
    (3, 6, 6, 61, 64, "f16", 7, True, "sdpa_mma"),                # MHA, D = 64: prologue + mma.sync key groups, crosses a tile edge
    (2, 16, 2, 120, 256, "bf16", 12, True, "sdpa_mma"),           # Qwen3.5 geometry (rope on 64 of 256 features)
    (1, 8, 2, 700, 64, "bf16", 5, False, "sdpa_mma"),             # one sequence: key range split over CTAs + merge
    (2, 4, 4, 90, 256, "f16", 4, True, "decode_simt"),            # wide heads, one query head per kv head: CUDA cores
    (2, 4, 4, 1, 128, "bf16", 3, False, "decode_hmma_tma"),       # nearly empty cache
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"B{c[0]}H{c[1]}k{c[2]}S{c[3]}D{c[4]}{c[5]}")
def test_dynamic_position_equals_host_offset_step(case):
    B, Hq, Hkv, S0, D, dtype, steps, norm, kernel = case
    (ce, cd), _, _ = _caches(B, Hkv, S0, D, dtype, 10)
    rope = omx.nn.Rope(D if D < 256 else 64, False, 1e6, 1.0)
    qn = kn = None
    if norm:
        qn, kn = omx.nn.RmsNorm(randn((D,), dtype, 5).to(DEV), 1e-6), omx.nn.RmsNorm(randn((D,), dtype, 6).to(DEV), 1e-6)
    max_rows = S0 + steps
    cd.prepare_graph(max_rows, Hq)
    pos = torch.full((1,), S0, dtype=torch.int32, device=DEV)
    for t in range(steps):
        q = randn((B, 1, Hq, D), dtype, 100 + 3 * t).transpose(1, 2).to(DEV)
        k = randn((B, 1, Hkv, D), dtype, 101 + 3 * t).transpose(1, 2).to(DEV)
        v = randn((B, 1, Hkv, D), dtype, 102 + 3 * t).transpose(1, 2).to(DEV)
        want = omx.attn_decode_fused(q, k, v, ce, rope, D ** -0.5, q_norm=qn, k_norm=kn)
        omx.launch_count(reset=True)
        got = omx.attn_decode_fused_dynamic(q, k, v, cd, rope, D ** -0.5, pos, q_norm=qn, k_norm=kn)
        assert omx.last_kernel() == kernel and (omx.launch_count() == 1 or kernel == "sdpa_mma")
        omx.device_counter_add(pos, 1)
        cd.advance(1)
        assert torch.equal(got, want), f"step {t}"
        assert cd.offset() == ce.offset() == S0 + t + 1
    assert int(pos.item()) == S0 + steps
    (ke, ve), (kd, vd) = ce.state(), cd.state()
    assert ke.shape == kd.shape  # logical capacity follows cache.rs:141-181 on both
    assert torch.equal(ke, kd) and torch.equal(ve, vd)
    # and the ordinary API keeps working on the pinned cache
    k = randn((B, Hkv, 2, D), dtype, 999).to(DEV)
    K1, _ = ce.update_and_fetch(k, k)
    K2, _ = cd.update_and_fetch(k, k)
    assert torch.equal(K1, K2)


def test_graph_replay_three_layers_vs_eager_and_oracle():
    B, Hq, Hkv, S0, D, dtype, L, steps = 2, 16, 4, 250, 128, "bf16", 3, 12
    rope = omx.nn.Rope(D, False, 1e6, 1.0)
    eager, graphed, ocache = [], [], None
    for layer in range(L):
        (ce, cg), k0, v0 = _caches(B, Hkv, S0, D, dtype, 20 + 2 * layer)
        eager.append(ce)
        graphed.append(cg)
        if layer == L - 1:
            ocache = orc.KVCache()
            ocache.update_and_fetch(t2n(k0, dtype), t2n(v0, dtype))
    qn = omx.nn.RmsNorm(randn((D,), dtype, 7).to(DEV), 1e-6)
    kn = omx.nn.RmsNorm(randn((D,), dtype, 8).to(DEV), 1e-6)
    qs = [torch.empty((B, 1, Hq, D), dtype=torch.bfloat16, device=DEV).transpose(1, 2) for _ in range(L)]
    ks = [torch.empty((B, 1, Hkv, D), dtype=torch.bfloat16, device=DEV).transpose(1, 2) for _ in range(L)]
    vs = [torch.empty((B, 1, Hkv, D), dtype=torch.bfloat16, device=DEV).transpose(1, 2) for _ in range(L)]
    loop = omx.DecodeLoopGraph(qs, ks, vs, graphed, rope, D ** -0.5, S0 + steps, q_norm=qn, k_norm=kn)
    assert loop.launches_per_step == L + 1
    assert all(c.offset() == S0 for c in graphed)  # the warm-up left no trace in the bookkeeping
    for t in range(steps):
        ins = []
        for layer in range(L):
            q = randn((B, 1, Hq, D), dtype, 1000 + 10 * t + layer).transpose(1, 2)
            k = randn((B, 1, Hkv, D), dtype, 2000 + 10 * t + layer).transpose(1, 2)
            v = randn((B, 1, Hkv, D), dtype, 3000 + 10 * t + layer).transpose(1, 2)
            qs[layer].copy_(q.to(DEV)); ks[layer].copy_(k.to(DEV)); vs[layer].copy_(v.to(DEV))
            ins.append((q, k, v))
        outs = loop.step()
        for layer in range(L):
            want = omx.attn_decode_fused(qs[layer], ks[layer], vs[layer], eager[layer], rope, D ** -0.5,
                                         q_norm=qn, k_norm=kn)
            assert torch.equal(outs[layer], want), f"step {t} layer {layer}"
        # oracle chain for the last layer
        q, k, v = (t2n(a, dtype) for a in ins[-1])
        off = ocache.offset()
        qo = orc.rope(orc.rms_norm(q, t2n(qn.weight, dtype), 1e-6, dtype=dtype), D, False, 1e6, 1.0, off, dtype=dtype)
        ko = orc.rope(orc.rms_norm(k, t2n(kn.weight, dtype), 1e-6, dtype=dtype), D, False, 1e6, 1.0, off, dtype=dtype)
        K, V = ocache.update_and_fetch(ko, v)
        o = orc.sdpa(qo, np.ascontiguousarray(K), np.ascontiguousarray(V), D ** -0.5, None, dtype=dtype)
        assert_close(outs[-1].float().cpu().numpy(), n2f(o, dtype), dtype, f"graph step {t} vs oracle")
    torch.cuda.synchronize()
    assert int(loop.position.item()) == S0 + steps
    for layer in range(L):
        assert graphed[layer].offset() == S0 + steps
        (ke, ve), (kg, vg) = eager[layer].state(), graphed[layer].state()
        assert ke.shape == kg.shape and torch.equal(ke, kg) and torch.equal(ve, vg)
    sk, sv = graphed[-1].state()
    assert_bits_equal(sk, ocache.keys, dtype, "KV cache keys after graph replays")
    assert_bits_equal(sv, ocache.values, dtype, "KV cache values after graph replays")
    with pytest.raises(omx.Exception, match="pinned for"):
        loop.step()


def test_graph_mode_errors_are_loud():
    B, Hq, Hkv, S0, D = 1, 8, 2, 10, 128
    (c,), _, _ = _caches(B, Hkv, S0, D, "bf16", 1, n=1)
    q = randn((B, Hq, 1, D), "bf16", 2).to(DEV)
    k = randn((B, Hkv, 1, D), "bf16", 3).to(DEV)
    pos = torch.full((1,), S0, dtype=torch.int32, device=DEV)
    rope = omx.nn.Rope(D, False, 1e6, 1.0)
    with pytest.raises(omx.Exception, match="prepare_graph"):
        omx.attn_decode_fused_dynamic(q, k, k, c, rope, 1.0, pos)
    with pytest.raises(omx.Exception, match="cache is empty"):
        omx.KVCache().prepare_graph(100, Hq)
    with pytest.raises(omx.Exception, match="leaves no room"):
        c.prepare_graph(S0, Hq)
    c.prepare_graph(S0 + 2, Hq)
    omx.attn_decode_fused_dynamic(q, k, k, c, rope, 1.0, pos)
    c.advance(2)
    with pytest.raises(omx.Exception, match="exceeds the"):
        c.advance(1)
    with pytest.raises(omx.Exception, match="int32"):
        omx.attn_decode_fused_dynamic(q, k, k, c, rope, 1.0, pos.long())
    # growing past the pinned rows through the ordinary API moves the buffers: graph mode must be re-armed
    big = randn((B, Hkv, 2000, D), "bf16", 4).to(DEV)
    c.update_and_fetch(big, big)
    with pytest.raises(omx.Exception, match="prepare_graph"):
        omx.attn_decode_fused_dynamic(q, k, k, c, rope, 1.0, pos)
