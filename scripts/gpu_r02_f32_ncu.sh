#!/bin/bash
cat > /tmp/f32probe.py <<'PY'
import importlib, os, sys
import torch
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
omx = importlib.import_module("ominix-mlx_b200")
q = torch.randn((1, 16, 2048, 128), device="cuda")
k = torch.randn((1, 8, 2048, 128), device="cuda")
v = torch.randn((1, 8, 2048, 128), device="cuda")
for _ in range(3):
    omx.fast.scaled_dot_product_attention(q, k, v, 128 ** -0.5, omx.fast.ScaledDotProductAttentionMask.Causal)
torch.cuda.synchronize()
PY
timeout 300 ncu --set full --import-source on --clock-control none -k regex:sdpa_f32_tiled -s 2 -c 1 -f -o gpurun_out/r02_f32_tiled python /tmp/f32probe.py > gpurun_out/r02_f32_ncu.log 2>&1
tail -2 gpurun_out/r02_f32_ncu.log
