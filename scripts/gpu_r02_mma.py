"""16-bit attention outside the tcgen05 / TMA kernels' shapes: the mma.sync kernel (sdpa_mma.cu) vs the row-per-warp
kernel (CUDA events after 3 warm-ups; decode shapes rotate through caches larger than L2)."""
import importlib, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
omx = importlib.import_module("ominix-mlx_b200")
Causal = omx.fast.ScaledDotProductAttentionMask.Causal
# name, B, Hq, Hkv, Lq, Lk, Dk, Dv, causal, generic reps
SHAPES = [("GLM-4.7-Flash absorbed MLA decode: 20 heads / 1 latent kv head, 576 / 512, B1, ctx 8192", 1, 20, 1, 1, 8192, 576, 512, False, 3),
          ("same, B16, ctx 4096", 16, 20, 1, 1, 4096, 576, 512, False, 2),
          ("same, B64, ctx 4096", 64, 20, 1, 1, 4096, 576, 512, False, 1),
          ("absorbed MLA prefill: B1, S2048, causal", 1, 20, 1, 2048, 2048, 576, 512, True, 1),
          ("absorbed MLA prefill: B1, S4096, causal", 1, 20, 1, 4096, 4096, 576, 512, True, 0),
          ("head_dim 256 prefill (qwen3.5): 16 q / 2 kv, B1, S4096, causal", 1, 16, 2, 4096, 4096, 256, 256, True, 1),
          ("head_dim 80 vision tower: 16 heads, B4, 1024 tokens, no mask", 4, 16, 16, 1024, 1024, 80, 80, False, 2),
          ("head_dim 72 vision tower: 16 heads, B4, 729 tokens, no mask", 4, 16, 16, 729, 729, 72, 72, False, 2)]
for name, B, Hq, Hkv, Lq, Lk, Dk, Dv, causal, greps in SHAPES:
    g = torch.Generator(device="cuda").manual_seed(1)
    kv_bytes = B * Hkv * Lk * (Dk + Dv) * 2
    nrot = max(1, min(16, int(300e6 // kv_bytes))) if Lq == 1 else 1
    q = torch.randn((B, Hq, Lq, Dk), generator=g, device="cuda").bfloat16()
    ks = [torch.randn((B, Hkv, Lk, Dk), generator=g, device="cuda").bfloat16() for _ in range(nrot)]
    vs = [torch.randn((B, Hkv, Lk, Dv), generator=g, device="cuda").bfloat16() for _ in range(nrot)]
    out = torch.empty((B, Hq, Lq, Dv), device="cuda", dtype=torch.bfloat16)
    flops = 2.0 * B * Hq * Lq * Lk * (Dk + Dv) * (0.5 if causal else 1.0)
    res, outs = {}, {}
    variants = [("sdpa_mma", 20 * nrot), ("sdpa_generic", (greps if not os.environ.get("OMX_ATTN_LIB") else 0) * nrot)]
    if Lq == 1 and Dk == 576:
        variants.insert(1, ("sdpa_mma/no_key_groups", 20 * nrot))
    for kern, reps in variants:
        if reps == 0:
            continue
        os.environ.pop("OMX_MMA_NO_KS", None)
        if kern.endswith("no_key_groups"):
            os.environ["OMX_MMA_NO_KS"] = "1"
        omx.force_kernel(kern.split("/")[0])
        try:
            for i in range(3):
                omx.fast.scaled_dot_product_attention(q, ks[i % nrot], vs[i % nrot], Dk ** -0.5, Causal if causal else None, out=out)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            graph = None
            if Lq == 1 and kern != "sdpa_generic":  # a ctypes call costs ~30 us of host time: replay a graph of the calls
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    for i in range(2):  # the library's per-stream scratch is allocated outside the capture
                        omx.fast.scaled_dot_product_attention(q, ks[i % nrot], vs[i % nrot], Dk ** -0.5, None, out=out)
                    side.synchronize()
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph, stream=side):
                        for i in range(reps):
                            omx.fast.scaled_dot_product_attention(q, ks[i % nrot], vs[i % nrot], Dk ** -0.5, None, out=out)
                torch.cuda.current_stream().wait_stream(side)
                graph.replay()
                torch.cuda.synchronize()
            e0.record()
            if graph is not None:
                graph.replay()
            else:
                for i in range(reps):
                    omx.fast.scaled_dot_product_attention(q, ks[i % nrot], vs[i % nrot], Dk ** -0.5, Causal if causal else None, out=out)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            res[kern] = {"ms": round(ms, 4), "TFLOP/s": round(flops / ms / 1e9, 2), "KV GB/s": round(kv_bytes / ms / 1e6, 1)}
            omx.fast.scaled_dot_product_attention(q, ks[0], vs[0], Dk ** -0.5, Causal if causal else None, out=out)
            outs[kern] = out.float().clone()
        finally:
            omx.force_kernel("")
            os.environ.pop("OMX_MMA_NO_KS", None)
    diff = float((outs["sdpa_mma"] - outs["sdpa_generic"]).abs().max()) if "sdpa_generic" in outs else None
    print(json.dumps({"shape": name, "rotated_caches": nrot, **res, "max_abs_diff_between_kernels": diff}), flush=True)
