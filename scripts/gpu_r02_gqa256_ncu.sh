#!/bin/bash
cat > /tmp/gqa_probe.py <<'PY'
import importlib, os, sys
import torch
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
omx = importlib.import_module("ominix-mlx_b200")
B, Hq, Hkv, S, D = 64, 16, 2, 8192, 256
q = torch.randn((B, Hq, 1, D), device="cuda").bfloat16()
k = torch.randn((B, Hkv, S, D), device="cuda").bfloat16()
v = torch.randn((B, Hkv, S, D), device="cuda").bfloat16()
for _ in range(3):
    omx.fast.scaled_dot_product_attention(q, k, v, D ** -0.5, None)
torch.cuda.synchronize()
PY
timeout 300 ncu --set full --import-source on --clock-control none -k regex:sdpa_mma_kernel -s 2 -c 1 -f -o gpurun_out/r02_gqa256 python /tmp/gqa_probe.py > gpurun_out/r02_gqa256_ncu.log 2>&1
tail -2 gpurun_out/r02_gqa256_ncu.log
