#!/bin/bash
# A/B in ONE call (boxes differ by several %): previous build vs current build, alternating.
for rep in 1 2; do
  for lib in libomx_attn_prev.so libomx_attn.so; do
    echo "== $lib (rep $rep)"
    OMX_ATTN_LIB=$PWD/ominix-mlx_b200/$lib timeout 300 python scripts/bench_small_decode.py 2>&1 | tail -4 | cut -c1-150
    echo "c2 $(OMX_ATTN_LIB=$PWD/ominix-mlx_b200/$lib timeout 300 python bench.py --steps 300 --warmup 10 --no-cpu 2>/dev/null | tail -1 | grep -o '"ms_per_step": [0-9.e-]*' | head -1)"
  done
done
