"""Single-sequence decode latency under a CUDA graph, KV rotated through R caches (cold L2):
fused step (rope + append + attention) vs plain sdpa (Lq = 1) on the same caches."""
import importlib, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
omx = importlib.import_module("ominix-mlx_b200")
dev = "cuda"
SHAPES = [("c1 fp32 16/8 ctx2048", 1, 16, 8, 2048, torch.float32), ("0.6b bf16 16/8 ctx2048", 1, 16, 8, 2048, torch.bfloat16),
          ("8b bf16 32/8 ctx8192", 1, 32, 8, 8192, torch.bfloat16), ("c5 bf16 32/8 ctx32768", 1, 32, 8, 32768, torch.bfloat16),
          ("c5 one rank of 8: bf16 4/1 ctx32768", 1, 4, 1, 32768, torch.bfloat16)]
R, D = 16, 128
for name, B, Hq, Hkv, S, dt in SHAPES:
    if len(sys.argv) > 1 and sys.argv[1] not in name:
        continue
    g = torch.Generator(device=dev).manual_seed(1)
    rn = lambda *s: torch.randn(s, generator=g, device=dev, dtype=torch.float32).to(dt)
    rope = omx.nn.Rope(D, False, 1e6, 1.0)
    qn, kn = omx.nn.RmsNorm(rn(D), 1e-6), omx.nn.RmsNorm(rn(D), 1e-6)
    k0, v0 = rn(B, Hkv, S - 1, D), rn(B, Hkv, S - 1, D)
    caches = []
    for _ in range(R):
        c = omx.KVCache(); c.reserve(S + 512); c.update_and_fetch(k0, v0); caches.append(c)
    q, k, v = rn(B, 1, Hq, D).transpose(1, 2), rn(B, 1, Hkv, D).transpose(1, 2), rn(B, 1, Hkv, D).transpose(1, 2)
    out = torch.empty((B, Hq, 1, D), dtype=dt, device=dev)
    views = [c.update_and_fetch(k, v) for c in caches]  # now S rows
    for c in caches:
        c.trim(1)
    res = {}
    def fused(norm):
        def f():
            for c in caches:
                omx.attn_decode_fused(q, k, v, c, rope, D ** -0.5, out=out, q_norm=qn if norm else None, k_norm=kn if norm else None)
                c.trim(1)
        return f
    def plain():
        for K, V in views:
            omx.fast.scaled_dot_product_attention(q, K, V, D ** -0.5, None, out=out)
    only = os.environ.get("OMX_BENCH_LABELS")
    for label, fn in (("fused", fused(False)), ("fused_norm", fused(True)), ("sdpa_only", plain)):
        if only and label not in only.split(","):
            continue
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            fn()
        for _ in range(5):
            gr.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            gr.replay()
        e1.record(); torch.cuda.synchronize()
        res[label] = round(1e3 * e0.elapsed_time(e1) / (50 * R), 2)
    es = 4 if dt == torch.float32 else 2
    res["floor_us_at_6534GBs"] = round(2 * B * Hkv * S * D * es / 6534e9 * 1e6, 2)
    print(json.dumps({"shape": name, "us_per_launch": res}), flush=True)
