"""Lq = 1 sdpa on 16-bit shapes outside head_dim 128: the CUDA-core split-K kernel (default dispatch) vs the mma.sync
kernel (forced), CUDA-graph replay over rotated K / V sets."""
import importlib, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
omx = importlib.import_module("ominix-mlx_b200")
dev, dt = "cuda", torch.bfloat16
SHAPES = [(32, 16, 2, 256, 4096), (32, 16, 4, 256, 4096), (32, 16, 16, 256, 2048), (1, 16, 2, 256, 8192),
          (32, 16, 4, 64, 4096), (32, 32, 4, 64, 4096), (32, 16, 16, 64, 4096), (1, 16, 4, 64, 8192),
          (32, 16, 4, 32, 4096), (32, 16, 4, 80, 4096), (32, 16, 4, 96, 4096)]
for B, Hq, Hkv, D, S in SHAPES:
    g = torch.Generator(device=dev).manual_seed(1)
    rn = lambda *s: torch.randn(s, generator=g, device=dev, dtype=torch.float32).to(dt)
    kv_bytes = 2 * B * Hkv * S * D * 2
    R = max(1, min(16, int(300e6 // kv_bytes)))
    ks, vs = [rn(B, Hkv, S, D) for _ in range(R)], [rn(B, Hkv, S, D) for _ in range(R)]
    q = rn(B, Hq, 1, D)
    out = torch.empty((B, Hq, 1, D), dtype=dt, device=dev)
    res = {}
    for label in ("default", "sdpa_mma"):
        def fn():
            for K, V in zip(ks, vs):
                omx.fast.scaled_dot_product_attention(q, K, V, D ** -0.5, None, out=out)
        try:
            omx.force_kernel("sdpa_mma" if label == "sdpa_mma" else "")
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    fn()
                kern = omx.last_kernel()
                side.synchronize()
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr, stream=side):
                    fn()
            torch.cuda.current_stream().wait_stream(side)
            for _ in range(3):
                gr.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                gr.replay()
            e1.record(); torch.cuda.synchronize()
            us = 1e3 * e0.elapsed_time(e1) / (20 * R)
            res[label] = {"kernel": kern, "us": round(us, 2), "TB/s": round(kv_bytes / us / 1e6, 2)}
            res[label + "_out"] = out.float().clone()
        except Exception as e:
            res[label] = {"error": str(e).splitlines()[0][:120]}
        finally:
            omx.force_kernel("")
    a, b = res.pop("default_out", None), res.pop("sdpa_mma_out", None)
    diff = float((a - b).abs().max()) if a is not None and b is not None else None
    print(json.dumps({"shape": f"B{B} {Hq}/{Hkv} heads d{D} ctx{S}", "rotated": R, **res, "max_abs_diff": diff}), flush=True)
