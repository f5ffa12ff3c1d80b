"""Head-sharded decode (BASELINE C5 layout) on the GPU: the peer-store output path of the decode
kernel on one device (world 1 through the raw C ABI), and, when the box has >= 2 GPUs, both
exchange spellings (NCCL all-gather baseline, fused peer stores) across 2 ranks against the
oracle's unsharded answer."""
import ctypes
import importlib
import os
import socket

import numpy as np
import pytest
import torch

from conftest import assert_bits_equal, assert_close, load_oracle, load_pkg, n2f, randn, t2n

pytestmark = pytest.mark.gpu
omx = load_pkg()
orc = load_oracle()
ROPE = (128, False, 1e6, 1.0)
DEV = "cuda"


def _oracle_full(k, v, q, kn, vn, dtype, S):
    oc = orc.KVCache()
    oc.update_and_fetch(t2n(k, dtype), t2n(v, dtype))
    qo = orc.rope(t2n(q, dtype), *ROPE, S, dtype=dtype)
    ko = orc.rope(t2n(kn, dtype), *ROPE, S, dtype=dtype)
    K, V = oc.update_and_fetch(ko, t2n(vn, dtype))
    want = orc.sdpa(qo, np.ascontiguousarray(K), np.ascontiguousarray(V), 128 ** -0.5, None, dtype=dtype)
    return n2f(want, dtype), oc


@pytest.mark.parametrize("dtype,S", [("bf16", 777), ("bf16", 5000), ("f32", 300)])
def test_peer_store_path_world1_raw_abi(dtype, S):
    """world = 1: the kernel's final store goes through the peer table (here: the local buffer,
    at a head offset inside a larger output) and the arrival counter reaches the step count."""
    L = omx._lib
    B, Hq, Hkv, D, Htot, h0 = 2, 8, 2, 128, 24, 8
    k, v = randn((B, Hkv, S, D), dtype, 1), randn((B, Hkv, S, D), dtype, 2)
    q, kn, vn = randn((B, Hq, 1, D), dtype, 3), randn((B, Hkv, 1, D), dtype, 4), randn((B, Hkv, 1, D), dtype, 5)
    want, oc = _oracle_full(k, v, q, kn, vn, dtype, S)
    cache = omx.KVCache()
    cache.update_and_fetch(k.cuda(), v.cuda())
    out_full = torch.full((B, Htot, 1, D), 7.0, dtype=q.dtype, device="cuda")
    flags = torch.zeros(8, dtype=torch.int32, device="cuda")
    pg = L.OmxPeerGroup()
    pg.world, pg.rank = 1, 0
    pg.out[0], pg.flags[0] = out_full.data_ptr(), flags.data_ptr()
    base = L.OmxOptionalFloat()
    base.has_value, base.value = True, ROPE[2]
    A = omx.array
    qg, kg, vg = q.cuda(), kn.cuda(), vn.cuda()  # descriptors borrow: keep the tensors alive
    qd, kd, vd, od = A.desc(qg), A.desc(kg), A.desc(vg), A.desc(out_full)
    sp = A.stream_ptr()
    for step in (1, 2):
        L.check(L.lib().omx_attn_decode_fused_sharded(A.ref(od), A.ref(qd), A.ref(kd), A.ref(vd), cache.handle,
                                                      128, False, base, 1.0, None, 128 ** -0.5,
                                                      ctypes.byref(pg), h0, sp))
        L.check(L.lib().omx_peer_wait(ctypes.byref(pg), step, sp))
        torch.cuda.synchronize()
        assert int(flags[0]) == step
        assert_close(out_full[:, h0:h0 + Hq].float().cpu().numpy(), want, dtype, "peer-store output slice")
        assert bool((out_full[:, :h0] == 7).all()) and bool((out_full[:, h0 + Hq:] == 7).all()), \
            "rows outside this rank's head slice were touched"
        if step == 1:
            sk, sv = cache.state()
            assert_bits_equal(sk, oc.keys, dtype, "KV keys (sharded step)")
            assert_bits_equal(sv, oc.values, dtype, "KV values (sharded step)")
            cache.trim(1)
    # validation: a head slice that does not fit
    assert L.lib().omx_attn_decode_fused_sharded(A.ref(od), A.ref(qd), A.ref(kd), A.ref(vd), cache.handle, 128,
                                                 False, base, 1.0, None, 0.1, ctypes.byref(pg), Htot - 2, sp) == 1
    assert b"do not fit" in L.lib().omx_last_error()


@pytest.mark.parametrize("dtype,B,Hq,Hkv,S", [("bf16", 1, 4, 1, 5000), ("bf16", 2, 8, 2, 777), ("f32", 1, 8, 2, 300),
                                              ("bf16", 1, 4, 1, 40)])
def test_ll_exchange_world1_raw_abi(dtype, B, Hq, Hkv, S):
    """world = 1 of the data + flag exchange (omx_attn_decode_fused_sharded_ll): the all-CTA combine's word path
    (long single sequence), the exchange as a second kernel (fp32 / short contexts), the step counter, and the
    private out_full untouched outside this rank's heads."""
    L = omx._lib
    D = 128
    k, v = randn((B, Hkv, S, D), dtype, 1), randn((B, Hkv, S, D), dtype, 2)
    q, kn, vn = randn((B, Hq, 1, D), dtype, 3), randn((B, Hkv, 1, D), dtype, 4), randn((B, Hkv, 1, D), dtype, 5)
    want, oc = _oracle_full(k, v, q, kn, vn, dtype, S)
    cache = omx.KVCache()
    cache.update_and_fetch(k.cuda(), v.cuda())
    out_full = torch.full((B, Hq, 1, D), 7.0, dtype=q.dtype, device="cuda")
    dt = L.OMX_FLOAT32 if dtype == "f32" else L.OMX_BFLOAT16
    nbytes = int(L.lib().omx_ll_staging_bytes(1, B, Hq, D, dt))
    assert nbytes == 2 * B * Hq * D * (4 if dtype == "f32" else 2) // 4 * 8
    staging = torch.zeros(nbytes // 8, dtype=torch.int64, device="cuda")
    seq = torch.zeros(1, dtype=torch.int32, device="cuda")
    g = L.OmxLLGroup()
    g.world, g.rank = 1, 0
    g.staging[0], g.seq = staging.data_ptr(), seq.data_ptr()
    base = L.OmxOptionalFloat()
    base.has_value, base.value = True, ROPE[2]
    A = omx.array
    qg, kg, vg = q.cuda(), kn.cuda(), vn.cuda()
    qd, kd, vd, od = A.desc(qg), A.desc(kg), A.desc(vg), A.desc(out_full)
    sp = A.stream_ptr()
    for step in (1, 2, 3):
        L.check(L.lib().omx_attn_decode_fused_sharded_ll(A.ref(od), A.ref(qd), A.ref(kd), A.ref(vd), cache.handle,
                                                         128, False, base, 1.0, None, 128 ** -0.5,
                                                         ctypes.byref(g), 0, sp))
        torch.cuda.synchronize()
        assert int(seq[0]) == step
        assert_close(out_full.float().cpu().numpy(), want, dtype, "ll exchange, world 1")
        if step == 1:
            sk, sv = cache.state()
            assert_bits_equal(sk, oc.keys, dtype, "KV keys (ll step)")
            assert_bits_equal(sv, oc.values, dtype, "KV values (ll step)")
        cache.trim(1)
        out_full.fill_(7.0)
    # validation: the head slice must be rank * Hq_local
    assert L.lib().omx_attn_decode_fused_sharded_ll(A.ref(od), A.ref(qd), A.ref(kd), A.ref(vd), cache.handle, 128,
                                                    False, base, 1.0, None, 0.1, ctypes.byref(g), 4, sp) == 1
    assert b"must write heads" in L.lib().omx_last_error()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rank_main(rank, world, port, gather, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        dtype, B, Hq, Hkv, D, S = "bf16", 1, 32, 8, 128, 4099
        k, v = randn((B, Hkv, S, D), dtype, 1), randn((B, Hkv, S, D), dtype, 2)
        qq, kn, vn = randn((B, Hq, 1, D), dtype, 3), randn((B, Hkv, 1, D), dtype, 4), randn((B, Hkv, 1, D), dtype, 5)
        want, oc = _oracle_full(k, v, qq, kn, vn, dtype, S)
        rope = omx.nn.Rope(*ROPE)
        eng = omx.parallel.HeadShardedDecode(Hq, Hkv, D, torch.bfloat16, rope, D ** -0.5, batch=B, gather=gather)
        eng.prefill(k.cuda(), v.cuda())
        errs = []
        for _ in range(3):
            out = eng.step(qq.cuda(), kn.cuda(), vn.cuda())
            torch.cuda.synchronize()
            errs.append(float(np.abs(out.float().cpu().numpy() - want).max()))
            eng.rewind(1)
        # back-to-back steps with NO barrier and one rank that is slow to read its output: the output buffer is
        # double-buffered by step parity, so the fast rank's next step must not overwrite what the slow rank is
        # still reading (write-after-read across ranks).  Distinct queries per step make a stale / torn read visible.
        qs = [qq, randn((B, Hq, 1, D), dtype, 13), randn((B, Hq, 1, D), dtype, 23)]
        wants = [want] + [_oracle_full(k, v, qx, kn, vn, dtype, S)[0] for qx in qs[1:]]
        qs_d, kn_d, vn_d = [x.cuda() for x in qs], kn.cuda(), vn.cuda()
        got = []
        for t in range(9):
            out = eng.step(qs_d[t % 3], kn_d, vn_d)
            if rank == 1:
                torch.cuda._sleep(3_000_000)  # ~1.5 ms on the stream before this rank consumes its output
            got.append(out.clone())
            eng.rewind(1)
        torch.cuda.synchronize()
        for t, g in enumerate(got):
            errs.append(float(np.abs(g.float().cpu().numpy() - wants[t % 3]).max()))
        sk, _ = eng.cache.state()
        kv_ok = bool((t2n(sk, dtype) == oc.keys[:, eng.kv0:eng.kv0 + eng.nkv]).all())
        q.put((rank, max(errs), kv_ok, omx.last_kernel()))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("gather", ["collective", "peer", "peer_flags"])
def test_head_sharded_decode_two_ranks(gather):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, gather, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0, f"rank process exited with {p.exitcode}"
    res = sorted(q.get(timeout=5) for _ in range(2))
    for rank, err, kv_ok, kern in res:
        assert err <= 2e-2, (rank, err)
        assert kv_ok, f"rank {rank}: KV shard not bit-exact"
        assert kern == "decode_hmma_tma"


# ---- sequence-sharded decode ----

def _seq_oracle(k, v, steps, dtype, S, Hq, Hkv, D, B):
    """Unsharded chain: rope(q, pos), rope(k_new, pos), append, sdpa -- per step."""
    oc = orc.KVCache()
    oc.update_and_fetch(t2n(k, dtype), t2n(v, dtype))
    outs = []
    for t, (q, kn, vn) in enumerate(steps):
        qo = orc.rope(t2n(q, dtype), D, False, 1e6, 1.0, S + t, dtype=dtype)
        ko = orc.rope(t2n(kn, dtype), D, False, 1e6, 1.0, S + t, dtype=dtype)
        K, V = oc.update_and_fetch(ko, t2n(vn, dtype))
        outs.append(n2f(orc.sdpa(qo, np.ascontiguousarray(K), np.ascontiguousarray(V), D ** -0.5, None, dtype=dtype), dtype))
    return outs, oc


@pytest.mark.parametrize("dtype,world", [("bf16", 1), ("bf16", 2), ("f32", 3), ("bf16", 4)])
def test_seq_sharded_virtual_ranks_one_gpu(dtype, world):
    """All `world` ranks played by ONE GPU through the raw C ABI: rank r's partial buffer and counters are plain
    device tensors, every peer pointer is a local address -- the kernels cannot tell.  Checks the partial / merge
    math, the global-position rope, the ownership rule and the (bit-exact) distributed cache rows."""
    import ctypes
    L = importlib.import_module("ominix-mlx_b200._lib")
    A = importlib.import_module("ominix-mlx_b200.array")
    B, Hq, Hkv, D, S, T = 2, 8, 2, 128, 203, 5
    tdt = torch.bfloat16 if dtype == "bf16" else torch.float32
    k, v = randn((B, Hkv, S, D), dtype, 1), randn((B, Hkv, S, D), dtype, 2)
    steps = [(randn((B, Hq, 1, D), dtype, 10 + 3 * t), randn((B, Hkv, 1, D), dtype, 11 + 3 * t),
              randn((B, Hkv, 1, D), dtype, 12 + 3 * t)) for t in range(T)]
    want, oc = _seq_oracle(k, v, steps, dtype, S, Hq, Hkv, D, B)
    caches = [omx.KVCache() for _ in range(world)]
    kd, vd = k.to(DEV), v.to(DEV)
    for r, c in enumerate(caches):
        rows = omx.parallel.seq_shard_rows(S, world, r).to(DEV)
        c.update_and_fetch(kd[:, :, rows], vd[:, :, rows])
    parts = [torch.zeros((2, world, B, Hq, D + 2), dtype=torch.float32, device=DEV) for _ in range(world)]
    flags = [torch.zeros(L.OMX_MAX_PEERS, dtype=torch.int32, device=DEV) for _ in range(world)]
    outs = [torch.empty((B, Hq, 1, D), dtype=tdt, device=DEV) for _ in range(world)]
    base = L.OmxOptionalFloat()
    base.has_value, base.value = True, 1e6
    sp = A.stream_ptr(None)
    for t, (q, kn, vn) in enumerate(steps):
        par_ = t & 1
        pos = S + t
        groups = []
        for r in range(world):
            pg = L.OmxPeerGroup()
            pg.world, pg.rank = world, r
            for j in range(world):
                pg.out[j] = parts[j][par_].data_ptr()
                pg.flags[j] = flags[j].data_ptr()
            groups.append(pg)
        qd, knd, vnd = q.to(DEV), kn.to(DEV), vn.to(DEV)
        for r in range(world):
            own = omx.parallel.seq_shard_owner(pos, world) == r
            L.check(L.lib().omx_attn_decode_seqshard(
                A.ref(A.desc(parts[r][par_])), A.ref(A.desc(qd)), A.ref(A.desc(knd)) if own else None,
                A.ref(A.desc(vnd)) if own else None, caches[r].handle, D, False, base, 1.0, pos, own, D ** -0.5,
                ctypes.byref(groups[r]), sp))
        for r in range(world):
            L.check(L.lib().omx_seqshard_merge(A.ref(A.desc(outs[r])), A.ref(A.desc(parts[r][par_])),
                                               ctypes.byref(groups[r]), ctypes.c_uint32(t + 1), sp))
        torch.cuda.synchronize()
        for r in range(world):
            assert_close(outs[r].float().cpu().numpy(), want[t], dtype, f"seq-sharded step {t} rank {r}")
            assert torch.equal(outs[r], outs[0])
        # the merge is what parallel.merge_partials says
        ref_merge = omx.parallel.merge_partials(parts[0][par_])
        np.testing.assert_allclose(outs[0][:, :, 0].float().cpu().numpy(), ref_merge.cpu().numpy(), rtol=1e-2, atol=1e-2)
    # every rank holds exactly the rows it owns, bit-exact
    for r, c in enumerate(caches):
        rows = omx.parallel.seq_shard_rows(S + T, world, r).numpy()
        assert c.offset() == len(rows)
        sk, sv = c.state()
        assert (t2n(sk[:, :, :len(rows)], dtype) == oc.keys[:, :, rows]).all()
        assert (t2n(sv[:, :, :len(rows)], dtype) == oc.values[:, :, rows]).all()


def _seq_rank_main(rank, world, port, gather, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        dtype, B, Hq, Hkv, D, S, T = "bf16", 1, 32, 8, 128, 4099, 4
        k, v = randn((B, Hkv, S, D), dtype, 1), randn((B, Hkv, S, D), dtype, 2)
        steps = [(randn((B, Hq, 1, D), dtype, 10 + 3 * t), randn((B, Hkv, 1, D), dtype, 11 + 3 * t),
                  randn((B, Hkv, 1, D), dtype, 12 + 3 * t)) for t in range(T)]
        want, oc = _seq_oracle(k, v, steps, dtype, S, Hq, Hkv, D, B)
        eng = omx.parallel.SeqShardedDecode(Hq, Hkv, D, torch.bfloat16, omx.nn.Rope(*ROPE), D ** -0.5, batch=B,
                                            gather=gather)
        eng.prefill(k.cuda(), v.cuda())
        errs = []
        for t, (qq, kn, vn) in enumerate(steps):
            out = eng.step(qq.cuda(), kn.cuda(), vn.cuda())
            torch.cuda.synchronize()
            errs.append(float(np.abs(out.float().cpu().numpy() - want[t]).max()))
        rows = omx.parallel.seq_shard_rows(S + T, world, rank).numpy()
        sk, _ = eng.cache.state()
        kv_ok = bool((t2n(sk[:, :, :len(rows)], dtype) == oc.keys[:, :, rows]).all())
        q.put((rank, max(errs), kv_ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("gather", ["collective", "peer"])
def test_seq_sharded_decode_two_ranks(gather):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_seq_rank_main, args=(r, 2, port, gather, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0, f"rank process exited with {p.exitcode}"
    for rank, err, kv_ok in sorted(q.get(timeout=5) for _ in range(2)):
        assert err <= 2e-2, (rank, err)
        assert kv_ok, f"rank {rank}: KV rows not bit-exact"


def test_auto_expected_wait_is_replayable_one_gpu():
    """expected == 0 ("as many arrivals as this rank has itself signalled") lets the sharded step -- decode launch +
    wait / merge -- be captured ONCE and replayed: two virtual ranks on one GPU, sequence-sharded, the captured
    pair of steps replayed three times against freshly computed eager results."""
    import ctypes
    L = importlib.import_module("ominix-mlx_b200._lib")
    A = importlib.import_module("ominix-mlx_b200.array")
    world, B, Hq, Hkv, D, S, dtype = 2, 1, 8, 2, 128, 301, "bf16"
    k, v = randn((B, Hkv, S, D), dtype, 1).to(DEV), randn((B, Hkv, S, D), dtype, 2).to(DEV)
    q, kn, vn = (randn(s, dtype, i).to(DEV) for i, s in ((3, (B, Hq, 1, D)), (4, (B, Hkv, 1, D)), (5, (B, Hkv, 1, D))))
    caches = [omx.KVCache() for _ in range(world)]
    for r, c in enumerate(caches):
        rows = omx.parallel.seq_shard_rows(S, world, r).to(DEV)
        c.update_and_fetch(k[:, :, rows], v[:, :, rows])
        c.reserve(1024)
    parts = [torch.zeros((world, B, Hq, D + 2), dtype=torch.float32, device=DEV) for _ in range(world)]
    flags = [torch.zeros(L.OMX_MAX_PEERS, dtype=torch.int32, device=DEV) for _ in range(world)]
    outs = [torch.empty((B, Hq, 1, D), dtype=torch.bfloat16, device=DEV) for _ in range(world)]
    base = L.OmxOptionalFloat()
    base.has_value, base.value = True, 1e6
    groups = []
    for r in range(world):
        pg = L.OmxPeerGroup()
        pg.world, pg.rank = world, r
        for j in range(world):
            pg.out[j], pg.flags[j] = parts[j].data_ptr(), flags[j].data_ptr()
        groups.append(pg)
    owner = omx.parallel.seq_shard_owner(S, world)

    def step(expected):
        sp = A.stream_ptr(None)
        for r in range(world):
            own = r == owner
            L.check(L.lib().omx_attn_decode_seqshard(
                A.ref(A.desc(parts[r])), A.ref(A.desc(q)), A.ref(A.desc(kn)) if own else None,
                A.ref(A.desc(vn)) if own else None, caches[r].handle, D, False, base, 1.0, S, own, D ** -0.5,
                ctypes.byref(groups[r]), sp))
        for r in range(world):
            L.check(L.lib().omx_seqshard_merge(A.ref(A.desc(outs[r])), A.ref(A.desc(parts[r])), ctypes.byref(groups[r]),
                                               ctypes.c_uint32(expected), sp))
        caches[owner].trim(1)  # identical work every step

    step(1)  # eager, explicit step count
    torch.cuda.synchronize()
    want = outs[0].clone()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        step(0)  # warm the capture stream
    side.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        step(0)
        step(0)
    for _ in range(3):
        for o in outs:
            o.zero_()
        g.replay()
        torch.cuda.synchronize()
        assert all(torch.equal(o, want) for o in outs)
    assert int(flags[0][0].item()) == 2 + 2 * 3  # every launch signalled: counters kept growing under replay
