"""One shape per process for an `ncu --set full` capture of sdpa_mma: argv[1] in {decode, prefill}."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
omx = importlib.import_module("ominix-mlx_b200")
if sys.argv[1] == "decode":
    B, L, S, mask = 64, 1, 4096, None
else:
    B, L, S, mask = 1, 2048, 2048, omx.fast.ScaledDotProductAttentionMask.Causal
q = torch.randn((B, 20, L, 576), device="cuda").bfloat16()
k = torch.randn((B, 1, S, 576), device="cuda").bfloat16()
v = torch.randn((B, 1, S, 512), device="cuda").bfloat16()
for _ in range(3):
    omx.fast.scaled_dot_product_attention(q, k, v, 576 ** -0.5, mask)
torch.cuda.synchronize()
