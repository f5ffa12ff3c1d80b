"""C3-shape prefill composite (q_norm / k_norm + rope + cache rows + causal attention) for an ncu capture of the
prologue kernel; prints CUDA-event times of composite vs plain attention."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
omx = importlib.import_module("ominix-mlx_b200")
B, Hq, Hkv, S, D = 8, 32, 8, 8192, 128
g = torch.Generator(device="cuda").manual_seed(3)
dt = torch.bfloat16
# the callers' layout: [B, L, H, D] storage viewed [B, H, L, D]
q = torch.randn((B, S, Hq, D), generator=g, device="cuda").to(dt).transpose(1, 2)
k = torch.randn((B, S, Hkv, D), generator=g, device="cuda").to(dt).transpose(1, 2)
v = torch.randn((B, S, Hkv, D), generator=g, device="cuda").to(dt).transpose(1, 2)
rope = omx.nn.Rope(D, False, 1e6, 1.0)
qn = omx.nn.RmsNorm(torch.ones(D, device="cuda", dtype=dt), 1e-6) if hasattr(omx.nn, "RmsNorm") else None
out = torch.empty((B, Hq, S, D), device="cuda", dtype=dt)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
cache = omx.KVCache()
cache.reserve(S)
def comp():
    cache.reset()  # offset back to 0: the same rows are written again (no allocation inside the timed loop)
    omx.attn_prefill_fused(q, k, v, cache, rope, D ** -0.5, out=out, q_norm=qn, k_norm=qn)
for _ in range(2):
    comp()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    comp()
e1.record()
torch.cuda.synchronize()
print("composite ms", e0.elapsed_time(e1) / reps, "kernel", omx.last_kernel())
qc, kc, vc = q.contiguous(), k.contiguous(), v.contiguous()
Causal = omx.fast.ScaledDotProductAttentionMask.Causal
for _ in range(2):
    omx.fast.scaled_dot_product_attention(qc, kc, vc, D ** -0.5, Causal, out=out)
e0.record()
for _ in range(reps):
    omx.fast.scaled_dot_product_attention(qc, kc, vc, D ** -0.5, Causal, out=out)
e1.record()
torch.cuda.synchronize()
print("plain attention ms", e0.elapsed_time(e1) / reps)
