"""Host-side sharding logic of SURVEY 8(e), covered on CPU with world_size-2 gloo.

The compute cannot run here (no CPU path in the product), so every rank plays its GPU's part with
the oracle -- the checker -- on ITS shard only, and the product's exchange code
(parallel.all_gather_heads, the shard arithmetic) must reassemble exactly what the oracle gives
for the unsharded problem."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import load_oracle, load_pkg, n2t, t2n

omx = load_pkg()
par = omx.parallel


def test_batch_shard_partitions_every_row_once():
    for batch in (0, 1, 3, 4, 8, 64, 65):
        for world in (1, 2, 4, 8):
            rows = []
            for r in range(world):
                s, n = par.batch_shard(batch, world, r)
                rows += list(range(s, s + n))
            assert rows == list(range(batch)), (batch, world)
            counts = [par.batch_shard(batch, world, r)[1] for r in range(world)]
            assert max(counts) - min(counts) <= 1
    with pytest.raises(omx.Exception):
        par.batch_shard(8, 2, 2)


def test_kv_head_shard_keeps_gqa_groups_together():
    # C5: 32 q / 8 kv heads over 8 ranks -> 1 kv head + its 4 q heads per rank
    for world in (1, 2, 4, 8):
        seen_q, seen_kv = [], []
        for r in range(world):
            kv0, nkv, q0, nq = par.kv_head_shard(32, 8, world, r)
            assert nq == nkv * 4 and q0 == kv0 * 4
            for h in range(q0, q0 + nq):  # every local q head reads a local kv head
                assert kv0 <= h // 4 < kv0 + nkv
            seen_q += list(range(q0, q0 + nq))
            seen_kv += list(range(kv0, kv0 + nkv))
        assert seen_q == list(range(32)) and seen_kv == list(range(8))
    with pytest.raises(omx.Exception, match="not divisible"):
        par.kv_head_shard(32, 8, 3, 0)
    with pytest.raises(omx.Exception, match="multiple"):
        par.kv_head_shard(30, 8, 2, 0)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, dtype, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        orc = load_oracle()
        orc.set_threads(1)
        B, Hq, Hkv, S, D = 1, 8, 4, 97, 64
        rng = np.random.default_rng(5)  # same stream on every rank = replicated inputs

        def mk(*shape):
            x = rng.standard_normal(shape, dtype=np.float32)
            return orc.f32_to_bf16_bits(x) if dtype == "bf16" else x

        q, k, v = mk(B, Hq, 1, D), mk(B, Hkv, S, D), mk(B, Hkv, S, D)
        scale = D ** -0.5
        full = orc.sdpa(q, k, v, scale, None, dtype=dtype)
        kv0, nkv, q0, nq = par.kv_head_shard(Hq, Hkv, world, rank)
        local = orc.sdpa(np.ascontiguousarray(q[:, q0:q0 + nq]), np.ascontiguousarray(k[:, kv0:kv0 + nkv]),
                         np.ascontiguousarray(v[:, kv0:kv0 + nkv]), scale, None, dtype=dtype)
        got = par.all_gather_heads(n2t(local, dtype), Hq)
        assert tuple(got.shape) == (B, Hq, 1, D)
        same = bool((t2n(got, dtype).view(np.uint16 if dtype == "bf16" else np.uint32) ==
                     np.ascontiguousarray(full).view(np.uint16 if dtype == "bf16" else np.uint32)).all())
        # batch sharding (C2/C3/C4): concatenating the per-rank results in rank order is the answer
        Bb = 5
        qb, kb, vb = mk(Bb, Hq, 1, D), mk(Bb, Hkv, S, D), mk(Bb, Hkv, S, D)
        fullb = orc.sdpa(qb, kb, vb, scale, None, dtype=dtype)
        s0, n = par.batch_shard(Bb, world, rank)
        mine = orc.sdpa(np.ascontiguousarray(qb[s0:s0 + n]), np.ascontiguousarray(kb[s0:s0 + n]),
                        np.ascontiguousarray(vb[s0:s0 + n]), scale, None, dtype=dtype)
        same_b = bool((np.ascontiguousarray(mine).view(np.uint8) ==
                       np.ascontiguousarray(fullb[s0:s0 + n]).view(np.uint8)).all())
        flags = torch.tensor([int(same), int(same_b)])
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        if rank == 0:
            ret.put(flags.tolist())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_head_sharded_gather_world2_gloo(dtype):
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, dtype, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret.get(timeout=5) == [1, 1]


# ---- sequence sharding: ownership + the log-sum-exp merge, world 2 over gloo ----

def test_seq_shard_rows_partition_every_position_once():
    for world in (1, 2, 3, 8):
        for start, n in ((0, 0), (0, 1), (0, 17), (5, 9), (7, 64)):
            seen = sorted(int(p) for r in range(world) for p in par.seq_shard_rows(n, world, r, start))
            assert seen == list(range(start, start + n)), (world, start, n)
            for r in range(world):
                assert all(par.seq_shard_owner(int(p), world) == r for p in par.seq_shard_rows(n, world, r, start))


def _partial(q, k, v, scale):
    """What omx_attn_decode_seqshard leaves in a slot: normalised output of the local keys + (m, l), log2 domain."""
    G = q.shape[1] // k.shape[1]
    kk, vv = k.repeat_interleave(G, 1), v.repeat_interleave(G, 1)
    s = torch.einsum("bhd,bhjd->bhj", q[:, :, 0].double(), kk.double()) * scale * 1.4426950408889634
    m = s.max(-1).values
    p = torch.exp2(s - m[..., None])
    l = p.sum(-1)
    o = torch.einsum("bhj,bhjd->bhd", p, vv.double()) / l[..., None]
    return torch.cat([o, m[..., None], l[..., None]], -1).float()


def _seq_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        B, Hq, Hkv, S, D = 2, 8, 2, 45, 32
        g = torch.Generator().manual_seed(11)  # replicated inputs
        q, k, v = (torch.randn(s, generator=g) for s in ((B, Hq, 1, D), (B, Hkv, S, D), (B, Hkv, S, D)))
        scale = D ** -0.5
        rows = par.seq_shard_rows(S, world, rank)
        mine = _partial(q, k[:, :, rows], v[:, :, rows], scale)
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        got = par.merge_partials(torch.stack(parts))
        kk, vv = k.repeat_interleave(Hq // Hkv, 1), v.repeat_interleave(Hq // Hkv, 1)
        want = torch.nn.functional.scaled_dot_product_attention(q.double(), kk.double(), vv.double(), scale=scale)[:, :, 0]
        ok = torch.allclose(got.double(), want, rtol=1e-5, atol=1e-6)
        # a rank without keys contributes (m, l) = (-inf, 0) and garbage output: it must not poison the merge
        empty = torch.full_like(mine, float("nan"))
        empty[..., -2], empty[..., -1] = float("-inf"), 0.0
        got2 = par.merge_partials(torch.stack(parts + [empty]))
        ok2 = torch.allclose(got2.double(), want, rtol=1e-5, atol=1e-6)
        flags = torch.tensor([int(ok), int(ok2)])
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        if rank == 0:
            ret.put(flags.tolist())
    finally:
        dist.destroy_process_group()


def test_seq_sharded_merge_world2_gloo():
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_seq_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret.get(timeout=5) == [1, 1]
