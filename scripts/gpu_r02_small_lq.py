"""Small query blocks (1 < Lq < 32) against a long cache: tcgen05 FMHA vs the CUDA-core generic kernel."""
import importlib, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
omx = importlib.import_module("ominix-mlx_b200")
dev = "cuda"
B, Hq, Hkv, D, S = 4, 32, 8, 128, 8192
g = torch.Generator(device=dev).manual_seed(1)
rn = lambda *s: torch.randn(s, generator=g, device=dev, dtype=torch.float32).bfloat16()
k, v = rn(B, Hkv, S, D), rn(B, Hkv, S, D)
for Lq in (2, 4, 8, 16, 31, 32, 64):
    q = rn(B, Hq, Lq, D)
    out = torch.empty_like(q)
    res = {}
    for kern in ("fmha_tcgen05", "sdpa_generic"):
        omx.force_kernel(kern)
        f = lambda: omx.fast.scaled_dot_product_attention(q, k, v, D ** -0.5, "causal", out=out)
        for _ in range(3):
            f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            f()
        e1.record(); torch.cuda.synchronize()
        res[kern] = round(e0.elapsed_time(e1) / 20 * 1e3, 1)
        res[kern + "_out"] = out.float().clone()
    omx.force_kernel("")
    err = float((res.pop("fmha_tcgen05_out") - res.pop("sdpa_generic_out")).abs().max())
    print(json.dumps({"Lq": Lq, "us": res, "max_abs_diff": err}), flush=True)
