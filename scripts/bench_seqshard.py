"""Sequence-sharded single-sequence decode on N GPUs: step time of
omx_attn_decode_seqshard + omx_seqshard_merge with the partials pushed over NVLink peer stores vs all-gathered by
NCCL, next to the kv-head-sharded layout (C5) on the same shape.  Launch:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_seqshard.py
Device time (CUDA events on the launching stream), max over ranks; one JSON line from rank 0."""
import importlib
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
omx = importlib.import_module("ominix-mlx_b200")


def timed(step, rewind, steps, warmup, dev):
    for _ in range(warmup):
        step(); rewind()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step(); rewind()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    B, Hq, Hkv, D, S = 1, 32, 8, 128, int(os.environ.get("CTX", 32768))
    g = torch.Generator(device=dev).manual_seed(7)  # same stream on every rank: replicated inputs
    rn = lambda *s: torch.randn(s, generator=g, device=dev, dtype=torch.float32).bfloat16()
    k, v = rn(B, Hkv, S - 1, D), rn(B, Hkv, S - 1, D)
    q, kn, vn = rn(B, Hq, 1, D), rn(B, Hkv, 1, D), rn(B, Hkv, 1, D)
    rope = omx.nn.Rope(D, False, 1e6, 1.0)
    res = {}
    outs = {}
    for gather in ("peer", "collective"):
        eng = omx.parallel.SeqShardedDecode(Hq, Hkv, D, torch.bfloat16, rope, D ** -0.5, batch=B, gather=gather)
        eng.prefill(k, v)

        def rewind():
            eng.position -= 1
            if omx.parallel.seq_shard_owner(eng.position, world) == rank:
                eng.cache.trim(1)
        res["seq_" + gather] = timed(lambda: eng.step(q, kn, vn), rewind, 300, 20, dev)
        outs["seq_" + gather] = eng.step(q, kn, vn).float().clone(); rewind()
        del eng
    if Hkv % world == 0:
        for gather in ("peer", "collective"):
            eng = omx.parallel.HeadShardedDecode(Hq, Hkv, D, torch.bfloat16, rope, D ** -0.5, batch=B, gather=gather)
            eng.prefill(k, v)
            res["head_" + gather] = timed(lambda: eng.step(q, kn, vn), lambda: eng.rewind(1), 300, 20, dev)
            outs["head_" + gather] = eng.step(q, kn, vn).float().clone(); eng.rewind(1)
            dist.barrier()
            del eng
    ref = outs["seq_collective"]
    diff = {n: float((o - ref).abs().max()) for n, o in outs.items()}
    if rank == 0:
        print(json.dumps({"bench": "single-sequence decode, sharded", "n_gpus": world, "ctx": S, "shape": "32 q / 8 kv heads, d 128, bf16",
                          "ms_per_step_max_over_ranks": res, "max_abs_diff_vs_seq_collective": diff,
                          "kv_bytes_per_gpu": 2 * Hkv * S * D * 2 // world}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
