"""C4 (FLUX.2-klein double-stream block: 512 txt + 4096 img tokens, 24 heads, d128, B4, bf16): the fused block attention
(norm q / k per stream + table rope + joint [txt; img] buffers in one prologue launch, then attention) vs attention alone."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
omx = importlib.import_module("ominix-mlx_b200")
B, H, D, St, Si = 4, 24, 128, 512, 4096
dt = torch.bfloat16
g = torch.Generator(device="cuda").manual_seed(4)
mk = lambda S: torch.randn((B, S, H, D), generator=g, device="cuda").to(dt)
qs, ks, vs = [mk(St), mk(Si)], [mk(St), mk(Si)], [mk(St), mk(Si)]
ang = torch.rand((B, St + Si, D // 2), generator=g, device="cuda") * 6.28
cos, sin = torch.cos(ang).to(dt), torch.sin(ang).to(dt)
norm = lambda: omx.nn.RmsNorm((1 + 0.1 * torch.randn(D, device="cuda")).to(dt), 1e-6)
qn, kn = [norm(), norm()], [norm(), norm()]
def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
omx.launch_count(reset=True)
omx.dit.attn_fused(qs, ks, vs, D ** -0.5, cos, sin, qn, kn)
print("fused block launches", omx.launch_count(), "kernel", omx.last_kernel())
print("fused block ms", round(timeit(lambda: omx.dit.attn_fused(qs, ks, vs, D ** -0.5, cos, sin, qn, kn)), 4))
q = torch.cat(qs, 1).transpose(1, 2).contiguous(); k = torch.cat(ks, 1).transpose(1, 2).contiguous(); v = torch.cat(vs, 1).transpose(1, 2).contiguous()
out = torch.empty_like(q)
print("attention alone ms", round(timeit(lambda: omx.fast.scaled_dot_product_attention(q, k, v, D ** -0.5, None, out=out)), 4))
