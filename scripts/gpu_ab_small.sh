#!/bin/bash
# A/B in ONE call (boxes differ by several %): previous build (libomx_attn_prev.so) vs current build.
for lib in libomx_attn_prev.so libomx_attn.so; do
  echo "== $lib"
  OMX_ATTN_LIB=$PWD/ominix-mlx_b200/$lib timeout 300 python scripts/bench_small_decode.py 2>&1 | tail -5 | cut -c1-170
  echo "c2 $(OMX_ATTN_LIB=$PWD/ominix-mlx_b200/$lib timeout 300 python bench.py --steps 300 --warmup 10 --no-cpu 2>/dev/null | tail -1 | grep -o '"ms_per_step": [0-9.e-]*' | head -1)"
done
timeout 300 python -m pytest tests/test_decode_gpu.py tests/test_decode_random_gpu.py tests/test_parallel_gpu.py tests/test_graph_decode_gpu.py -x -q 2>&1 | tail -2
