"""Debug driver: smallest hmma decode call (non-fused then fused), prints max error."""
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
omx = importlib.import_module("ominix-mlx_b200")
from oracle import oracle as orc
B, Hq, Hkv, S, D = int(os.environ.get("B", 1)), int(os.environ.get("HQ", 4)), int(os.environ.get("HKV", 1)), int(os.environ.get("S", 128)), 128
g = torch.Generator().manual_seed(0)
def rn(*s): return torch.randn(s, generator=g).bfloat16()
def bits(t): return t.contiguous().cpu().view(torch.int16).numpy().view(np.uint16).copy()
q, k, v = rn(B, Hq, 1, D), rn(B, Hkv, S, D), rn(B, Hkv, S, D)
out = omx.fast.scaled_dot_product_attention(q.cuda(), k.cuda(), v.cuda(), D ** -0.5)
torch.cuda.synchronize()
print("kernel", omx.last_kernel())
want = orc.bf16_bits_to_f32(orc.sdpa(bits(q), bits(k), bits(v), D ** -0.5, None, dtype="bf16"))
print("non-fused max err", np.abs(out.float().cpu().numpy() - want).max())
if os.environ.get("FUSED", "1") == "1":
    c = omx.KVCache(); c.update_and_fetch(k.cuda(), v.cuda())
    kn, vn = rn(B, Hkv, 1, D), rn(B, Hkv, 1, D)
    o2 = omx.attn_decode_fused(q.cuda(), kn.cuda(), vn.cuda(), c, omx.nn.Rope(D, False, 1e6, 1.0), D ** -0.5)
    torch.cuda.synchronize()
    print("fused ok", omx.last_kernel(), o2.float().abs().max().item())
