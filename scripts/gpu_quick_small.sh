#!/bin/bash
timeout 600 python -m pytest tests/test_decode_gpu.py tests/test_graph_decode_gpu.py tests/test_golden_gpu.py tests/test_parallel_gpu.py -x -q 2>&1 | tail -3
timeout 300 python scripts/bench_small_decode.py 2>&1 | tail -4
for wl in c2 c5 c1; do
  echo "$wl $(timeout 300 python bench.py --workload $wl --steps 200 --warmup 10 --no-cpu --graph --rotate 16 2>/dev/null | tail -1 | grep -o '"ms_per_step": [0-9.e-]*' | head -1)"
done
