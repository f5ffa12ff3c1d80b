"""Error of the tcgen05 FMHA vs the oracle (bf16 chain) and vs float64, many short causal rows."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import importlib
import numpy as np, torch
omx = importlib.import_module("ominix-mlx_b200")
from oracle import oracle as orc
sys.path.insert(0, "tests")
from conftest import randn, t2n, n2f
for D in (64, 128):
    for seed in (0, 1, 2):
        B, H, L = 2, 24, 400
        q, k, v = (randn((B, H, L, D), "bf16", seed * 10 + s) for s in (1, 2, 3))
        got = omx.fast.scaled_dot_product_attention(q.cuda(), k.cuda(), v.cuda(), D ** -0.5, "causal").float().cpu().numpy()
        want = n2f(orc.sdpa(t2n(q, "bf16"), t2n(k, "bf16"), t2n(v, "bf16"), D ** -0.5, "causal", dtype="bf16"), "bf16")
        ex = torch.nn.functional.scaled_dot_product_attention(q.double(), k.double(), v.double(), is_causal=True).numpy()
        e1 = np.abs(got - want); e2 = np.abs(got - ex); e3 = np.abs(want - ex)
        print(f"D{D} seed{seed} kernel={omx.last_kernel()} got-oracle max {e1.max():.4f} (n>2e-2: {(e1 > 2e-2).sum()}) "
              f"got-exact max {e2.max():.4f} mean {e2.mean():.2e} | oracle-exact max {e3.max():.4f} mean {e3.mean():.2e}")
