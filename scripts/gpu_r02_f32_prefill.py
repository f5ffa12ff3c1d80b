"""float32 prefill: the tiled FFMA kernel vs the row-per-warp kernel (CUDA events, 20 calls after 3 warm-ups)."""
import importlib, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
omx = importlib.import_module("ominix-mlx_b200")
Causal = omx.fast.ScaledDotProductAttentionMask.Causal
SHAPES = [("C1 model prefill: 16 q / 8 kv, d128, B1, S2048, causal", 1, 16, 8, 2048, 2048, 128, True),
          ("same, S8192", 1, 16, 8, 8192, 8192, 128, True),
          ("FLUX f32 example: 24 heads, d128, B1, 4608 tokens, no mask", 1, 24, 24, 4608, 4608, 128, False),
          ("head_dim 64: 16 heads, B2, S2048, causal", 2, 16, 16, 2048, 2048, 64, True)]
for name, B, Hq, Hkv, Lq, Lk, D, causal in SHAPES:
    g = torch.Generator(device="cuda").manual_seed(1)
    q = torch.randn((B, Hq, Lq, D), generator=g, device="cuda")
    k = torch.randn((B, Hkv, Lk, D), generator=g, device="cuda")
    v = torch.randn((B, Hkv, Lk, D), generator=g, device="cuda")
    out = torch.empty_like(q)
    flops = 4.0 * B * Hq * Lq * Lk * D * (0.5 if causal else 1.0)
    res = {}
    for kern, reps in (("sdpa_f32_tiled", 20), ("sdpa_generic", 3)):
        omx.force_kernel(kern)
        try:
            for _ in range(2):
                omx.fast.scaled_dot_product_attention(q, k, v, D ** -0.5, Causal if causal else None, out=out)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                omx.fast.scaled_dot_product_attention(q, k, v, D ** -0.5, Causal if causal else None, out=out)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            res[kern] = {"ms": round(ms, 3), "TFLOP/s": round(flops / ms / 1e9, 2)}
            res[kern + "_out"] = out.clone()
        finally:
            omx.force_kernel("")
    diff = float((res.pop("sdpa_f32_tiled_out") - res.pop("sdpa_generic_out")).abs().max())
    print(json.dumps({"shape": name, **res, "max_abs_diff_between_kernels": diff}), flush=True)
