#!/bin/bash
# key tiles per split (OMX_MMA_MIN_TILES, read once per process) against single-sequence latencies
for mt in 1 2 4 8; do
  echo "== OMX_MMA_MIN_TILES=$mt"
  OMX_MMA_MIN_TILES=$mt timeout 120 python - <<'PY'
import importlib, os, sys, json
import torch
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
omx = importlib.import_module("ominix-mlx_b200")
dt = torch.bfloat16
for name, B, Hq, Hkv, S, Dk, Dv in (("MLA B1 ctx8192", 1, 20, 1, 8192, 576, 512), ("MLA B1 ctx32768", 1, 20, 1, 32768, 576, 512),
                                    ("d256 16/2 B1 ctx8192", 1, 16, 2, 8192, 256, 256), ("d64 16/4 B1 ctx8192", 1, 16, 4, 8192, 64, 64),
                                    ("MLA B4 ctx8192", 4, 20, 1, 8192, 576, 512)):
    R = max(1, min(16, int(300e6 // (B * Hkv * S * (Dk + Dv) * 2))))
    q = torch.randn((B, Hq, 1, Dk), device="cuda").to(dt)
    ks = [torch.randn((B, Hkv, S, Dk), device="cuda").to(dt) for _ in range(R)]
    vs = [torch.randn((B, Hkv, S, Dv), device="cuda").to(dt) for _ in range(R)]
    out = torch.empty((B, Hq, 1, Dv), device="cuda", dtype=dt)
    def fn():
        for K, V in zip(ks, vs):
            omx.fast.scaled_dot_product_attention(q, K, V, Dk ** -0.5, None, out=out)
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn(); fn(); side.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): g.replay()
    e1.record(); torch.cuda.synchronize()
    print("  %-24s %7.2f us  (%s)" % (name, 1e3 * e0.elapsed_time(e1) / (20 * R), omx.last_kernel()))
PY
done
