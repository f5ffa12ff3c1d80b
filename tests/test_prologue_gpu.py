"""The one-launch q/k/v prologue (csrc/prologue.cu): RMSNorm -> rope -> KV-cache rows / joint DiT buffers.
Bit-exact against the oracle's op chain (same bits as the standalone kernels), and the composites that
use it: omx_attn_prefill_fused (2 launches) and omx_dit_attn_fused (klein_model.rs:443-489, :641-663)."""
import numpy as np
import pytest
import torch

from conftest import assert_bits_equal, assert_close, load_oracle, load_pkg, n2f, n2t, randn, t2n, tdt

pytestmark = pytest.mark.gpu
omx = load_pkg()
orc = load_oracle()
DEV = "cuda"
Causal = omx.fast.ScaledDotProductAttentionMask.Causal


def _weights(D, dtype, seed):
    return (1 + 0.1 * randn((D,), "f32", seed)).to(tdt(dtype))


@pytest.mark.parametrize("dtype", ["bf16", "f16", "f32"])
@pytest.mark.parametrize("D,rope_t", [(128, (128, False, 1e6, 1.0)), (128, (128, True, 1e6, 1.0)),
                                      (128, (64, True, 10000.0, 1.0)), (128, (64, False, 10000.0, 0.5)),
                                      (64, (64, False, 10000.0, 1.0)), (128, None)])
@pytest.mark.parametrize("norms", [False, True])
def test_prefill_prologue_is_one_launch_and_cache_bit_exact(dtype, D, rope_t, norms):
    B, Hq, Hkv, L = 2, 8, 2, 70
    gc, oc = omx.KVCache(), orc.KVCache()
    for step, Lc in enumerate((L, 33)):  # second chunk starts at a non-zero position
        q = randn((B, Lc, Hq, D), dtype, 10 * step + 1).transpose(1, 2)
        k = randn((B, Lc, Hkv, D), dtype, 10 * step + 2).transpose(1, 2)
        v = randn((B, Lc, Hkv, D), dtype, 10 * step + 3).transpose(1, 2)
        qw = kw = qn = kn = None
        if norms:
            qw, kw = _weights(D, dtype, 4), _weights(D, dtype, 5)
            qn, kn = omx.nn.RmsNorm(qw.to(DEV), 1e-6), omx.nn.RmsNorm(kw.to(DEV), 1e-6)
        rope = None if rope_t is None else omx.nn.Rope(*rope_t)
        qd, kd, vd = q.to(DEV), k.to(DEV), v.to(DEV)
        omx.launch_count(reset=True)
        got = omx.attn_prefill_fused(qd, kd, vd, gc, rope, D ** -0.5, Causal, q_norm=qn, k_norm=kn)
        n_launch = omx.launch_count()
        # prologue + attention (the cache may add a grow/zero-fill on the first chunk: not counted as launches)
        assert n_launch == 2, f"expected prologue + attention, got {n_launch} launches"
        qo, ko = t2n(q, dtype), t2n(k, dtype)
        if norms:
            qo = orc.rms_norm(qo, t2n(qw, dtype), 1e-6, dtype=dtype)
            ko = orc.rms_norm(ko, t2n(kw, dtype), 1e-6, dtype=dtype)
        off = oc.offset()
        if rope_t is not None:
            qo = orc.rope(qo, *rope_t, off, dtype=dtype)
            ko = orc.rope(ko, *rope_t, off, dtype=dtype)
        K, V = oc.update_and_fetch(ko, t2n(v, dtype))
        want = orc.sdpa(qo, np.ascontiguousarray(K), np.ascontiguousarray(V), D ** -0.5, "causal", dtype=dtype)
        sk, sv = gc.state()
        assert_bits_equal(sk, oc.keys, dtype, "KV keys written by the prologue")
        assert_bits_equal(sv, oc.values, dtype, "KV values written by the prologue")
        assert_close(got.float().cpu().numpy(), n2f(want, dtype), dtype, f"prefill after prologue, chunk {step}")


def test_prologue_equals_standalone_ops_bitwise_at_scale():
    # q' is internal to the composite: its bits are pinned by demanding the SAME attention output as the
    # unfused library chain (standalone rms_norm / rope kernels, each bit-exact vs the oracle elsewhere)
    B, Hq, Hkv, L, D = 2, 32, 8, 1024, 128
    q, k, v = (randn((B, L, h, D), "bf16", s).to(DEV).transpose(1, 2) for h, s in ((Hq, 1), (Hkv, 2), (Hkv, 3)))
    qn = omx.nn.RmsNorm(_weights(D, "bf16", 4).to(DEV), 1e-6)
    kn = omx.nn.RmsNorm(_weights(D, "bf16", 5).to(DEV), 1e-6)
    rope = omx.nn.Rope(D, False, 1e6, 1.0)
    c1, c2 = omx.KVCache(), omx.KVCache()
    a = omx.attn_prefill_fused(q, k, v, c1, rope, D ** -0.5, Causal, q_norm=qn, k_norm=kn)
    b = omx.attn_decode_unfused(q, k, v, c2, rope, D ** -0.5, q_norm=qn, k_norm=kn)
    assert torch.equal(a, b)
    assert torch.equal(c1.state()[0], c2.state()[0]) and torch.equal(c1.state()[1], c2.state()[1])


def _dit_inputs(B, H, Hkv, D, lens, dtype, seed):
    qs = [randn((B, S, H, D), dtype, seed + 10 * i) for i, S in enumerate(lens)]
    ks = [randn((B, S, Hkv, D), dtype, seed + 10 * i + 1) for i, S in enumerate(lens)]
    vs = [randn((B, S, Hkv, D), dtype, seed + 10 * i + 2) for i, S in enumerate(lens)]
    S = sum(lens)
    axes = [32, 32, 32, 32] if D == 128 else [16, 16, 16, 16]
    ids = torch.randint(0, 64, (B, S, len(axes)), generator=torch.Generator().manual_seed(seed + 7)).float()
    c, s = orc.klein_rope_freqs(ids.numpy(), axes, 2000.0)
    return qs, ks, vs, torch.from_numpy(c).to(tdt(dtype)), torch.from_numpy(s).to(tdt(dtype))


@pytest.mark.parametrize("dtype", ["bf16", "f32", "f16"])
@pytest.mark.parametrize("lens", [(24, 104), (37,)])
def test_dit_attn_fused_vs_oracle_chain(dtype, lens):
    # double-stream ([txt, img]) and single-stream blocks of FLUX.2-klein: norm -> rope -> concat -> attention
    B, H, D = 2, 24, 128
    qs, ks, vs, ct, st = _dit_inputs(B, H, H, D, lens, dtype, 100)
    qw = [_weights(D, dtype, 50 + i) for i in range(len(lens))]
    kw = [_weights(D, dtype, 60 + i) for i in range(len(lens))]
    qn = [omx.nn.RmsNorm(w.to(DEV), 1e-6) for w in qw]
    kn = [omx.nn.RmsNorm(w.to(DEV), 1e-6) for w in kw]
    omx.launch_count(reset=True)
    got = omx.dit.attn_fused([t.to(DEV) for t in qs], [t.to(DEV) for t in ks], [t.to(DEV) for t in vs], D ** -0.5,
                             cos=ct.to(DEV), sin=st.to(DEV), q_norm=qn, k_norm=kn, out_dtype=torch.float32)
    assert omx.launch_count() == 2, "prologue + attention"
    # oracle: per-stream rms_norm and table rope (slices of the joint tables), concat, manual chain
    t0, Qo, Ko, Vo = 0, [], [], []
    for i, S in enumerate(lens):
        c_i, s_i = t2n(ct[:, t0:t0 + S], dtype), t2n(st[:, t0:t0 + S], dtype)
        qn_i = orc.rms_norm(t2n(qs[i], dtype), t2n(qw[i], dtype), 1e-6, dtype=dtype)
        kn_i = orc.rms_norm(t2n(ks[i], dtype), t2n(kw[i], dtype), 1e-6, dtype=dtype)
        Qo.append(orc.dit_rope(qn_i, c_i, s_i, dtype))
        Ko.append(orc.dit_rope(kn_i, c_i, s_i, dtype))
        Vo.append(t2n(vs[i], dtype))
        t0 += S
    tr = lambda parts: np.ascontiguousarray(np.swapaxes(np.concatenate(parts, axis=1), 1, 2))  # noqa: E731
    want = orc.dit_attention(tr(Qo), tr(Ko), tr(Vo), dtype, np.float32(np.sqrt(D)))
    assert_close(np.swapaxes(got.cpu().numpy(), 1, 2), want, dtype, f"dit fused {lens}")
    # and bit-identical to the library's own unfused spelling (standalone norm / rope kernels + concat)
    t0, Qg, Kg = 0, [], []
    for i, S in enumerate(lens):
        c_i, s_i = ct[:, t0:t0 + S].to(DEV), st[:, t0:t0 + S].to(DEV)
        Qg.append(omx.dit.apply_rope(qn[i](qs[i].to(DEV)), c_i, s_i))
        Kg.append(omx.dit.apply_rope(kn[i](ks[i].to(DEV)), c_i, s_i))
        assert_bits_equal(Qg[-1], Qo[i], dtype, "standalone norm + dit rope vs oracle")
        t0 += S
    ref = omx.dit.joint_attention(torch.cat(Qg, 1), torch.cat(Kg, 1), torch.cat([t.to(DEV) for t in vs], 1),
                                  D ** -0.5, out_dtype=torch.float32)
    assert torch.equal(got, ref)


def test_dit_attn_fused_zimage_gqa_mask_and_plain_paths():
    # Z-Image: fewer kv heads (repeat_axis GQA), additive mask, one stream (zimage_model.rs:345-388)
    B, H, Hkv, D, S = 1, 8, 4, 128, 40
    qs, ks, vs, ct, st = _dit_inputs(B, H, Hkv, D, (S,), "bf16", 200)
    m = torch.zeros(S, S)
    m[:, S - 6:] = float("-inf")
    qn, kn = omx.nn.RmsNorm(_weights(D, "bf16", 1).to(DEV), 1e-5), omx.nn.RmsNorm(_weights(D, "bf16", 2).to(DEV), 1e-5)
    got = omx.dit.attn_fused(qs[0].to(DEV), ks[0].to(DEV), vs[0].to(DEV), D ** -0.5, cos=ct.to(DEV), sin=st.to(DEV),
                             q_norm=qn, k_norm=kn, add_mask=m.to(DEV), out_dtype=torch.float32)
    Q = omx.dit.apply_rope(qn(qs[0].to(DEV)), ct.to(DEV), st.to(DEV))
    K = omx.dit.apply_rope(kn(ks[0].to(DEV)), ct.to(DEV), st.to(DEV))
    rep = lambda t: t.repeat_interleave(H // Hkv, dim=2)  # noqa: E731
    ref = omx.dit.joint_attention(Q, rep(K), rep(vs[0].to(DEV)), D ** -0.5, add_mask=m.to(DEV),
                                  out_dtype=torch.float32)
    assert (got - ref).abs().max().item() <= 1e-5
    # no norm, no rope, one stream: nothing to do in front of the attention kernel
    omx.launch_count(reset=True)
    o = omx.dit.attn_fused(qs[0].to(DEV), ks[0].to(DEV), vs[0].to(DEV), D ** -0.5)
    assert omx.launch_count() == 1
    assert torch.equal(o, omx.dit.joint_attention(qs[0].to(DEV), ks[0].to(DEV), vs[0].to(DEV), D ** -0.5))
    # head dim outside the one-launch kernel's coverage: composed from the standalone ops, same answer
    c80 = torch.rand(1, 36, 40)
    x80 = [randn((1, n, 4, 80), "f32", 5 + i) for i, n in enumerate((16, 20))]
    w80 = omx.nn.RmsNorm(_weights(80, "f32", 9).to(DEV), 1e-6)
    o80 = omx.dit.attn_fused([t.to(DEV) for t in x80], [t.to(DEV) for t in x80], [t.to(DEV) for t in x80], 80 ** -0.5,
                             cos=c80.to(DEV), sin=(1 - c80 * c80).sqrt().to(DEV), q_norm=w80, k_norm=w80)
    t0, parts = 0, []
    for t in x80:
        n = t.shape[1]
        parts.append(omx.dit.apply_rope(w80(t.to(DEV)), c80[:, t0:t0 + n].to(DEV),
                                        (1 - c80 * c80).sqrt()[:, t0:t0 + n].to(DEV)))
        t0 += n
    J = torch.cat(parts, 1)
    ref80 = omx.dit.joint_attention(J, J, torch.cat([t.to(DEV) for t in x80], 1), 80 ** -0.5)
    assert torch.equal(o80, ref80)


def test_dit_attn_fused_full_size_c4_equals_unfused():
    # BASELINE C4: 512 txt + 4096 img tokens, 24 heads, D128, batch 4, bf16
    B, H, D, lens = 4, 24, 128, (512, 4096)
    g = torch.Generator(device=DEV).manual_seed(4)
    mk = lambda S: torch.randn((B, S, H, D), generator=g, device=DEV).bfloat16()  # noqa: E731
    qs, ks, vs = [mk(S) for S in lens], [mk(S) for S in lens], [mk(S) for S in lens]
    S = sum(lens)
    ang = torch.rand((B, S, D // 2), generator=g, device=DEV) * 6.28
    ct, st = ang.cos().bfloat16(), ang.sin().bfloat16()
    w = [omx.nn.RmsNorm((1 + 0.1 * torch.randn(D, generator=g, device=DEV)).bfloat16(), 1e-6) for _ in range(4)]
    got = omx.dit.attn_fused(qs, ks, vs, D ** -0.5, cos=ct, sin=st, q_norm=w[:2], k_norm=w[2:])
    assert omx.last_kernel() == "fmha_tcgen05"
    t0, Q, K = 0, [], []
    for i, n in enumerate(lens):
        Q.append(omx.dit.apply_rope(w[i](qs[i]), ct[:, t0:t0 + n], st[:, t0:t0 + n]))
        K.append(omx.dit.apply_rope(w[2 + i](ks[i]), ct[:, t0:t0 + n], st[:, t0:t0 + n]))
        t0 += n
    ref = omx.dit.joint_attention(torch.cat(Q, 1), torch.cat(K, 1), torch.cat(vs, 1), D ** -0.5)
    assert torch.equal(got, ref)
    # rows of the softmax are convex weights: every output lies inside the range of V (size-independent)
    V = torch.cat(vs, 1).float()
    assert (got.float().amax(1) <= V.amax(1) + 2e-2).all() and (got.float().amin(1) >= V.amin(1) - 2e-2).all()
