"""Qwen3.5 full-attention prefill (16 / 2 heads, head dim 256, rope on 64 features, q / k norm): the composite
(prologue + attention) vs the attention alone, and the rows against the standalone ops (bit-exact)."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
omx = importlib.import_module("ominix-mlx_b200")
B, Hq, Hkv, S, D = 2, 16, 2, 4096, 256
dt = torch.bfloat16
g = torch.Generator(device="cuda").manual_seed(3)
q = torch.randn((B, S, Hq, D), generator=g, device="cuda").to(dt).transpose(1, 2)
k = torch.randn((B, S, Hkv, D), generator=g, device="cuda").to(dt).transpose(1, 2)
v = torch.randn((B, S, Hkv, D), generator=g, device="cuda").to(dt).transpose(1, 2)
rope = omx.nn.Rope(64, False, 1e7, 1.0)
qn = omx.nn.RmsNorm(1 + 0.1 * torch.randn(D, device="cuda").to(dt), 1e-6)
kn = omx.nn.RmsNorm(1 + 0.1 * torch.randn(D, device="cuda").to(dt), 1e-6)
out = torch.empty((B, Hq, S, D), device="cuda", dtype=dt)
cache = omx.KVCache(); cache.reserve(S)
def comp():
    cache.reset()
    omx.attn_prefill_fused(q, k, v, cache, rope, D ** -0.5, out=out, q_norm=qn, k_norm=kn)
def timeit(fn, reps=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
omx.launch_count(reset=True)
comp()
print("composite launches", omx.launch_count(), "kernel", omx.last_kernel(), "ms", round(timeit(comp), 4))
qc, kc, vc = q.contiguous(), k.contiguous(), v.contiguous()
Causal = omx.fast.ScaledDotProductAttentionMask.Causal
print("attention alone ms", round(timeit(lambda: omx.fast.scaled_dot_product_attention(qc, kc, vc, D ** -0.5, Causal, out=out)), 4))
# rows vs the standalone ops
comp()
K, V = cache.state()
kr = omx.fast.rope(kn.forward(k), 64, False, 1e7, 1.0, 0)
print("cache keys bit-exact vs rms_norm -> rope:", bool(torch.equal(K[:, :, :S].contiguous().view(torch.int16), kr.contiguous().view(torch.int16))),
      "values:", bool(torch.equal(V[:, :, :S].contiguous().view(torch.int16), v.contiguous().view(torch.int16))))
