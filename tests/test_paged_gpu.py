"""Paged KV cache (north_star: "fuses RoPE and the paged KV append") vs the oracle's KVCache
(mlx-rs-core/src/cache.rs:134-194): the rows a sequence holds, page by page, are BIT-EXACT the rows the
reference's rectangular cache holds; the fused paged decode step equals the oracle's op chain; growth never
moves a page; sequences of one batch may have different lengths (each checked against its own oracle cache)."""
import numpy as np
import pytest
import torch

from conftest import assert_bits_equal, assert_close, load_oracle, load_pkg, n2f, randn, t2n, tdt

pytestmark = pytest.mark.gpu
omx = load_pkg()
orc = load_oracle()
DEV = "cuda"
ROPE = (128, False, 1e6, 1.0)


def _oracle_step(ocache, q, k_new, v_new, dtype, rope, scale, qw=None, kw=None, eps=1e-6):
    if qw is not None:
        q = orc.rms_norm(q, qw, eps, dtype=dtype)
    if kw is not None:
        k_new = orc.rms_norm(k_new, kw, eps, dtype=dtype)
    off = ocache.offset()
    if rope is not None:
        q = orc.rope(q, *rope, off, dtype=dtype)
        k_new = orc.rope(k_new, *rope, off, dtype=dtype)
    K, V = ocache.update_and_fetch(k_new, v_new)
    return orc.sdpa(q, np.ascontiguousarray(K), np.ascontiguousarray(V), scale, None, dtype=dtype)


@pytest.mark.parametrize("dtype,D,kernel", [("bf16", 128, "decode_hmma_tma"), ("f32", 128, "decode_simt"),
                                            ("f16", 64, "sdpa_mma"), ("bf16", 256, "sdpa_mma")])
def test_lockstep_batch_matches_the_reference_cache(dtype, D, kernel):
    """Same call sequence as the reference's decode loop: prefill n rows, then single-token steps across two
    page boundaries.  Materialised K/V == oracle KVCache rows bit for bit; outputs within the bar."""
    B, Hq, Hkv, S0 = 3, 8, 2, 100
    pc = omx.PagedKVCache(B, Hkv, D, tdt(dtype), n_pages=B * 4, max_pages_per_seq=4)
    oc = orc.KVCache()
    k, v = randn((B, Hkv, S0, D), dtype, 1), randn((B, Hkv, S0, D), dtype, 2)
    gk, gv = pc.update_and_fetch(k.to(DEV), v.to(DEV))
    ok, ov = oc.update_and_fetch(t2n(k, dtype), t2n(v, dtype))
    assert_bits_equal(gk, ok, dtype, "paged prefill keys")
    assert_bits_equal(gv, ov, dtype, "paged prefill values")
    rd = D if D < 256 else 64  # head dim 256: Qwen3.5's partial rotary
    rope = omx.nn.Rope(rd, False, 1e6, 1.0)
    rope_t = (rd, False, 1e6, 1.0)
    for t in range(60):  # 100 -> 160 rows: crosses the 128-row page boundary
        q = randn((B, 1, Hq, D), dtype, 100 + t).transpose(1, 2)  # caller layout: [B,L,H,D] viewed [B,H,L,D]
        kn = randn((B, 1, Hkv, D), dtype, 200 + t).transpose(1, 2)
        vn = randn((B, 1, Hkv, D), dtype, 300 + t).transpose(1, 2)
        omx.launch_count(reset=True)
        got = omx.attn_decode_fused_paged(q.to(DEV), kn.to(DEV), vn.to(DEV), pc, rope, D ** -0.5)
        assert omx.last_kernel() == kernel and (omx.launch_count() == 1 or kernel == "sdpa_mma")
        want = _oracle_step(oc, t2n(q, dtype), t2n(kn, dtype), t2n(vn, dtype), dtype, rope_t, D ** -0.5)
        if t % 13 == 0 or t == 59:
            assert_close(got.float().cpu().numpy(), n2f(want, dtype), dtype, f"paged fused decode step {t}")
    assert pc.offset() == oc.offset() == 160 and pc.lengths() == [160] * B
    gk, gv = pc.fetch()
    assert_bits_equal(gk, oc.keys[:, :, :160], dtype, "paged keys after 60 fused appends")
    assert_bits_equal(gv, oc.values[:, :, :160], dtype, "paged values after 60 fused appends")


def test_c2_geometry_with_norms_bf16():
    # Qwen3-8B geometry (32 q / 8 kv heads, D 128) with q_norm / k_norm folded in, ctx 1000 -> split-K plan
    B, Hq, Hkv, D, S0, dtype = 2, 32, 8, 128, 1000, "bf16"
    pc = omx.PagedKVCache(B, Hkv, D, torch.bfloat16, n_pages=40, max_pages_per_seq=20)
    oc = orc.KVCache()
    k, v = randn((B, Hkv, S0, D), dtype, 5), randn((B, Hkv, S0, D), dtype, 6)
    pc.update_and_fetch(k.to(DEV), v.to(DEV), fetch=False)
    oc.update_and_fetch(t2n(k, dtype), t2n(v, dtype))
    qw, kw = (1 + 0.1 * randn((D,), "f32", 7)).to(torch.bfloat16), (1 + 0.1 * randn((D,), "f32", 8)).to(torch.bfloat16)
    qn, kn_ = omx.nn.RmsNorm(qw.to(DEV), 1e-6), omx.nn.RmsNorm(kw.to(DEV), 1e-6)
    rope = omx.nn.Rope(*ROPE)
    for t in range(3):
        q, kn, vn = randn((B, Hq, 1, D), dtype, 10 + t), randn((B, Hkv, 1, D), dtype, 20 + t), randn((B, Hkv, 1, D), dtype, 30 + t)
        got = omx.attn_decode_fused_paged(q.to(DEV), kn.to(DEV), vn.to(DEV), pc, rope, D ** -0.5, q_norm=qn, k_norm=kn_)
        want = _oracle_step(oc, t2n(q, dtype), t2n(kn, dtype), t2n(vn, dtype), dtype, ROPE, D ** -0.5,
                            t2n(qw, dtype), t2n(kw, dtype))
        assert_close(got.float().cpu().numpy(), n2f(want, dtype), dtype, f"paged fused decode + norms, step {t}")
    gk, gv = pc.fetch()
    assert_bits_equal(gk, oc.keys[:, :, :S0 + 3], dtype, "paged keys (norm + rope + append)")
    assert_bits_equal(gv, oc.values[:, :, :S0 + 3], dtype, "paged values")


@pytest.mark.parametrize("D,B,S0", [(128, 4, 777), (64, 4, 777), (256, 1, 3000)])
def test_paged_equals_contiguous_cache_bitwise(D, B, S0):
    # same kernels, same arithmetic: the paged step's output equals the contiguous cache's bit for bit (head dim 128:
    # the decode kernels; 64 / 256: prologue + mma.sync key groups, one sequence at 3000 rows = split + merge)
    Hq, Hkv = 16, 4
    k, v = randn((B, Hkv, S0, D), "bf16", 1, DEV), randn((B, Hkv, S0, D), "bf16", 2, DEV)
    pc = omx.PagedKVCache(B, Hkv, D, torch.bfloat16, n_pages=64, max_pages_per_seq=48 if B == 1 else 16)
    cc = omx.KVCache()
    pc.update_and_fetch(k, v, fetch=False)
    cc.update_and_fetch(k, v)
    rope = omx.nn.Rope(D if D < 256 else 64, False, 1e6, 1.0)
    for t in range(4):
        q, kn, vn = (randn((B, h, 1, D), "bf16", 50 + 3 * t + i, DEV) for i, h in enumerate((Hq, Hkv, Hkv)))
        a = omx.attn_decode_fused_paged(q, kn, vn, pc, rope, D ** -0.5)
        b = omx.attn_decode_fused(q, kn, vn, cc, rope, D ** -0.5)
        assert torch.equal(a, b), f"step {t}: paged and contiguous outputs differ"
    gk, gv = pc.fetch()
    ck, cv = cc.state()
    assert torch.equal(gk, ck[:, :, :S0 + 4]) and torch.equal(gv, cv[:, :, :S0 + 4])


@pytest.mark.parametrize("D", [128, 64])
def test_ragged_batch_release_and_reuse(D):
    """Sequences of different lengths in one launch; a released slot is skipped; a reset slot starts over and
    reuses freed pages.  Every sequence is checked against its OWN oracle cache.  (D = 64: the mma.sync route.)"""
    B, Hq, Hkv, dtype = 4, 8, 2, "bf16"
    ROPE = (D, False, 1e6, 1.0)
    lens0 = [5, 64, 130, 0]  # one empty sequence: its first step attends the new key only
    pc = omx.PagedKVCache(B, Hkv, D, torch.bfloat16, n_pages=12, max_pages_per_seq=6)
    ocs = [orc.KVCache() for _ in range(B)]
    for b, n in enumerate(lens0):
        if n:
            k, v = randn((1, Hkv, n, D), dtype, 10 + b), randn((1, Hkv, n, D), dtype, 20 + b)
            pc.append_slot(b, k.to(DEV), v.to(DEV))
            ocs[b].update_and_fetch(t2n(k, dtype), t2n(v, dtype))
    assert pc.lengths() == lens0 and pc.free_pages() == 12 - (1 + 1 + 3 + 0)
    rope = omx.nn.Rope(*ROPE)
    active = [True] * B

    def step(t):
        q, kn, vn = randn((B, Hq, 1, D), dtype, 100 + t), randn((B, Hkv, 1, D), dtype, 200 + t), randn((B, Hkv, 1, D), dtype, 300 + t)
        out = torch.full((B, Hq, 1, D), 7.0, dtype=torch.bfloat16, device=DEV)
        omx.attn_decode_fused_paged(q.to(DEV), kn.to(DEV), vn.to(DEV), pc, rope, D ** -0.5, out=out)
        for b in range(B):
            if not active[b]:
                assert bool((out[b] == 7).all()), f"released slot {b} was written"
                continue
            want = _oracle_step(ocs[b], t2n(q[b:b + 1], dtype), t2n(kn[b:b + 1], dtype), t2n(vn[b:b + 1], dtype), dtype,
                                ROPE, D ** -0.5)
            assert_close(out[b:b + 1].float().cpu().numpy(), n2f(want, dtype), dtype, f"ragged step {t} seq {b}")

    for t in range(3):
        step(t)
    assert pc.lengths() == [8, 67, 133, 3]
    pc.release(1)  # sequence 1 finished: its pages return to the pool, the slot is skipped
    active[1] = False
    assert pc.lengths()[1] == -1 and pc.free_pages() == 12 - (1 + 3 + 1)
    for t in range(3, 5):
        step(t)
    pc.reset(1)  # a new request takes the slot
    active[1] = True
    ocs[1] = orc.KVCache()
    k, v = randn((1, Hkv, 70, D), dtype, 77), randn((1, Hkv, 70, D), dtype, 78)
    pc.append_slot(1, k.to(DEV), v.to(DEV))
    ocs[1].update_and_fetch(t2n(k, dtype), t2n(v, dtype))
    for t in range(5, 8):
        step(t)
    gk, gv = pc.fetch()
    for b in range(B):
        n = ocs[b].offset()
        assert pc.lengths()[b] == n
        assert_bits_equal(gk[b:b + 1, :, :n], ocs[b].keys[:, :, :n], dtype, f"ragged keys seq {b}")
        assert_bits_equal(gv[b:b + 1, :, :n], ocs[b].values[:, :, :n], dtype, f"ragged values seq {b}")
        assert bool((gk[b, :, n:] == 0).all()), "rows past a sequence's end must read +0.0"


def test_growth_is_zero_copy_and_pool_exhaustion_is_loud():
    B, Hkv, D = 1, 1, 128
    pc = omx.PagedKVCache(B, Hkv, D, torch.bfloat16, n_pages=3, max_pages_per_seq=3)
    rope = omx.nn.Rope(*ROPE)
    k = randn((B, Hkv, 60, D), "bf16", 1, DEV)
    pc.update_and_fetch(k, k, fetch=False)
    kp0, vp0, table0 = pc.pages()
    q = randn((B, 4, 1, D), "bf16", 2, DEV)
    kn = randn((B, Hkv, 1, D), "bf16", 3, DEV)
    for _ in range(100):  # 60 -> 160 rows: two more pages
        omx.attn_decode_fused_paged(q, kn, kn, pc, rope, 0.1)
    kp1, vp1, table1 = pc.pages()
    assert (kp1, vp1) == (kp0, vp0), "the pool moved"
    assert table1[0][:1] == table0[0] and len(table1[0]) == 3 and len(set(table1[0])) == 3
    assert pc.free_pages() == 0
    for _ in range(32):  # fill the last page: 160 -> 192
        omx.attn_decode_fused_paged(q, kn, kn, pc, rope, 0.1)
    with pytest.raises(omx.Exception, match="pages"):
        omx.attn_decode_fused_paged(q, kn, kn, pc, rope, 0.1)
    assert pc.lengths() == [192], "a refused step must not advance the sequence"


def test_trim_and_graph_replay_with_reserved_pages():
    # reserve() pre-assigns pages so that the step needs no host-side allocation: two steps (one per length buffer)
    # captured into a CUDA graph, replayed, equal to eager stepping
    B, Hq, Hkv, D, S0 = 2, 8, 2, 128, 120
    k, v = randn((B, Hkv, S0, D), "bf16", 1, DEV), randn((B, Hkv, S0, D), "bf16", 2, DEV)
    q, kn, vn = (randn((B, h, 1, D), "bf16", 5 + i, DEV) for i, h in enumerate((Hq, Hkv, Hkv)))
    rope = omx.nn.Rope(*ROPE)
    pe = omx.PagedKVCache(B, Hkv, D, torch.bfloat16, n_pages=16, max_pages_per_seq=8)
    pe.update_and_fetch(k, v, fetch=False)
    eager = [omx.attn_decode_fused_paged(q, kn, vn, pe, rope, D ** -0.5).clone() for _ in range(6)]
    pg = omx.PagedKVCache(B, Hkv, D, torch.bfloat16, n_pages=16, max_pages_per_seq=8)
    pg.update_and_fetch(k, v, fetch=False)
    pg.reserve(16)
    out = torch.empty_like(eager[0])
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):  # warm-up on the capture stream, then rewind
        omx.attn_decode_fused_paged(q, kn, vn, pg, rope, D ** -0.5, out=out)
        omx.attn_decode_fused_paged(q, kn, vn, pg, rope, D ** -0.5, out=out)
        pg.trim(2)
    side.synchronize()
    assert pg.lengths() == [S0] * B
    outs = [torch.empty_like(out) for _ in range(2)]
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        omx.attn_decode_fused_paged(q, kn, vn, pg, rope, D ** -0.5, out=outs[0])
        omx.attn_decode_fused_paged(q, kn, vn, pg, rope, D ** -0.5, out=outs[1])
    pg.sync_lengths()  # capture advanced the host mirror but ran nothing: the device lengths are the truth
    assert pg.lengths() == [S0] * B
    for r in range(3):
        g.replay()
        torch.cuda.synchronize()
        assert torch.equal(outs[0], eager[2 * r]) and torch.equal(outs[1], eager[2 * r + 1]), f"replay {r}"
    pg.sync_lengths()
    assert pg.lengths() == [S0 + 6] * B
    gk, _ = pg.fetch()
    ek, _ = pe.fetch()
    assert torch.equal(gk[:, :, :S0 + 6], ek[:, :, :S0 + 6])
