"""A few eager calls of the MLA decode shapes for an ncu launch list (kernel vs combine durations)."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
omx = importlib.import_module("ominix-mlx_b200")
for B, S in ((1, 8192), (16, 4096), (64, 4096)):
    g = torch.Generator(device="cuda").manual_seed(1)
    q = torch.randn((B, 20, 1, 576), generator=g, device="cuda").bfloat16()
    k = torch.randn((B, 1, S, 576), generator=g, device="cuda").bfloat16()
    v = torch.randn((B, 1, S, 512), generator=g, device="cuda").bfloat16()
    for _ in range(3):
        omx.fast.scaled_dot_product_attention(q, k, v, 576 ** -0.5, None)
    torch.cuda.synchronize()
q = torch.randn((1, 20, 2048, 576), device="cuda").bfloat16()
k = torch.randn((1, 1, 2048, 576), device="cuda").bfloat16()
v = torch.randn((1, 1, 2048, 512), device="cuda").bfloat16()
for _ in range(2):
    omx.fast.scaled_dot_product_attention(q, k, v, 576 ** -0.5, omx.fast.ScaledDotProductAttentionMask.Causal)
torch.cuda.synchronize()
