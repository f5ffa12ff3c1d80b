"""Qwen3.5-35B-A3B full-attention layers (qwen3.5-35B-mlx/src/attention.rs: 16 q / 2 kv heads, head_dim 256, rope on the
first 64 features): fused decode step and plain sdpa (Lq = 1) under a CUDA graph, KV rotated through R caches."""
import importlib, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
omx = importlib.import_module("ominix-mlx_b200")
dev, dt, D, Hq, Hkv = "cuda", torch.bfloat16, 256, 16, 2
for B, S in ((1, 8192), (16, 8192), (64, 8192)):
    g = torch.Generator(device=dev).manual_seed(1)
    rn = lambda *s: torch.randn(s, generator=g, device=dev, dtype=torch.float32).to(dt)
    kv_bytes = 2 * B * Hkv * S * D * 2
    R = max(1, min(16, int(300e6 // kv_bytes)))
    rope = omx.nn.Rope(64, False, 1e7, 1.0)
    qn, kn = omx.nn.RmsNorm(rn(D), 1e-6), omx.nn.RmsNorm(rn(D), 1e-6)
    k0, v0 = rn(B, Hkv, S - 1, D), rn(B, Hkv, S - 1, D)
    caches = []
    for _ in range(R):
        c = omx.KVCache(); c.reserve(S + 512); c.update_and_fetch(k0, v0); caches.append(c)
    q, k, v = rn(B, 1, Hq, D).transpose(1, 2), rn(B, 1, Hkv, D).transpose(1, 2), rn(B, 1, Hkv, D).transpose(1, 2)
    out = torch.empty((B, Hq, 1, D), dtype=dt, device=dev)
    views = [c.update_and_fetch(k, v) for c in caches]
    for c in caches:
        c.trim(1)
    res, kern = {}, {}
    def fused():
        for c in caches:
            omx.attn_decode_fused(q, k, v, c, rope, D ** -0.5, out=out, q_norm=qn, k_norm=kn)
            c.trim(1)
    def plain():
        for K, V in views:
            omx.fast.scaled_dot_product_attention(q, K, V, D ** -0.5, None, out=out)
    def plain_mma():
        omx.force_kernel("sdpa_mma")
        try:
            plain()
        finally:
            omx.force_kernel("")
    for label, fn in (("fused_norm", fused), ("sdpa_only", plain), ("sdpa_only/sdpa_mma", plain_mma)):
      try:
        for _ in range(3):
            fn()
        kern[label] = omx.last_kernel()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record(); torch.cuda.synchronize()
        eager = 1e3 * e0.elapsed_time(e1) / (5 * R)
        res[label] = {"eager_us": round(eager, 2), "kernel": kern[label]}
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            fn()
            side.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=side):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        for _ in range(3):
            gr.replay()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            gr.replay()
        e1.record(); torch.cuda.synchronize()
        us = 1e3 * e0.elapsed_time(e1) / (20 * R)
        res[label].update({"us": round(us, 2), "TB/s": round(kv_bytes / us / 1e6, 2)})
      except Exception as e:
        res.setdefault(label, {})["error"] = str(e).splitlines()[0][:160]
    print(json.dumps({"shape": f"B{B} ctx{S} 16/2 heads d256 bf16", "rotated": R, **res,
                      "floor_us_at_6534GBs": round(kv_bytes / 6534e9 * 1e6, 1)}), flush=True)
