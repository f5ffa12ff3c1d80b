"""Decode LOOP timing (not the bench.py contract line): one token step = n_layers fused decode launches of one
model's attention, positions advancing every step -- eager launches through the C ABI vs one captured CUDA graph
replayed per token (omx.DecodeLoopGraph, omx_attn_decode_fused_dynamic).  Prints one JSON line per shape.

    python scripts/bench_decode_loop.py [--steps 200]

Shapes: Qwen3-0.6B (C1: 28 layers, 16q/8kv, fp32 and bf16, B1, ctx 2048), Qwen3-8B (36 layers, 32q/8kv bf16,
B1 ctx 8192 and B8 ctx 4096), Mixtral-8x7B (32 layers, 32q/8kv bf16, B1, ctx 32768)."""
import argparse
import importlib
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
omx = importlib.import_module("ominix-mlx_b200")

SHAPES = [
    ("qwen3-0.6b fp32 B1 ctx2048", 28, 1, 16, 8, 2048, torch.float32, True),
    ("qwen3-0.6b bf16 B1 ctx2048", 28, 1, 16, 8, 2048, torch.bfloat16, True),
    ("qwen3-8b bf16 B1 ctx8192", 36, 1, 32, 8, 8192, torch.bfloat16, True),
    ("qwen3-8b bf16 B8 ctx4096", 36, 8, 32, 8, 4096, torch.bfloat16, True),
    ("mixtral-8x7b bf16 B1 ctx32768", 32, 1, 32, 8, 32768, torch.bfloat16, False),
]


def run(name, L, B, Hq, Hkv, S, dt, norm, steps, warmup, modes=("eager", "graph")):
    D, dev = 128, "cuda"
    g = torch.Generator(device=dev).manual_seed(1)

    def rn(*shape):
        return torch.randn(shape, generator=g, device=dev, dtype=torch.float32).to(dt)
    rope = omx.nn.Rope(D, False, 1e6, 1.0)
    qn = omx.nn.RmsNorm(rn(D), 1e-6) if norm else None
    kn = omx.nn.RmsNorm(rn(D), 1e-6) if norm else None
    qs = [rn(B, 1, Hq, D).transpose(1, 2) for _ in range(L)]
    ks = [rn(B, 1, Hkv, D).transpose(1, 2) for _ in range(L)]
    vs = [rn(B, 1, Hkv, D).transpose(1, 2) for _ in range(L)]
    S0 = S - steps - warmup - 1
    k0, v0 = rn(B, Hkv, S0, D), rn(B, Hkv, S0, D)

    def caches():
        out = []
        for _ in range(L):
            c = omx.KVCache()
            c.reserve(S + 512)
            c.update_and_fetch(k0, v0)
            out.append(c)
        return out
    res = {}
    for mode in modes:
        cs = caches()
        outs = [torch.empty((B, Hq, 1, D), dtype=dt, device=dev) for _ in range(L)]
        if mode == "graph":
            loop = omx.DecodeLoopGraph(qs, ks, vs, cs, rope, D ** -0.5, S, q_norm=qn, k_norm=kn, out=outs)
            step = loop.step
        else:
            def step():
                for i in range(L):
                    omx.attn_decode_fused(qs[i], ks[i], vs[i], cs[i], rope, D ** -0.5, out=outs[i], q_norm=qn, k_norm=kn)
        for _ in range(warmup):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        res[mode] = {"ms_per_token_step": ms, "us_per_layer": 1e3 * ms / L, "tokens_per_s": B * 1e3 / ms,
                     "final_offset": cs[0].offset(), "checksum": float(sum(o.float().sum().item() for o in outs))}
        del cs
    es = torch.finfo(dt).bits // 8
    kv_bytes = 2 * B * Hkv * S * D * es * L
    if len(res) < 2:
        print(json.dumps({"shape": name, **res}), flush=True)
        return
    print(json.dumps({"shape": name, "layers": L, "batch": B, "ctx": S, "dtype": str(dt).split(".")[-1],
                      "kv_bytes_per_token_step": kv_bytes, "eager": res["eager"], "graph": res["graph"],
                      "graph_speedup": res["eager"]["ms_per_token_step"] / res["graph"]["ms_per_token_step"],
                      "graph_hbm_gbs": kv_bytes / (res["graph"]["ms_per_token_step"] * 1e-3) / 1e9,
                      "same_result": res["eager"]["checksum"] == res["graph"]["checksum"]}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--only", default="")
    ap.add_argument("--modes", default="eager,graph")
    a = ap.parse_args()
    for s in SHAPES:
        if a.only and a.only not in s[0]:
            continue
        run(*s, a.steps, a.warmup, tuple(a.modes.split(",")))


if __name__ == "__main__":
    main()
