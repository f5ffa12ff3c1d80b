#!/bin/bash
# Graph-mode decode: new tests, golden fixtures, loop timing, C2/C5 regression check.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_graph_decode_gpu.py tests/test_golden_gpu.py -x -q > gpurun_out/g_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/g_pytest.log
tail -30 gpurun_out/g_pytest.log
timeout 600 python scripts/bench_decode_loop.py --steps 200 > gpurun_out/g_loop.jsonl 2> gpurun_out/g_loop.err; tail -3 gpurun_out/g_loop.err; cat gpurun_out/g_loop.jsonl
for wl in c2 c5 c1; do
  timeout 300 python bench.py --workload $wl --steps 100 --warmup 10 --no-cpu --graph 2>/dev/null | tail -1 | cut -c1-260
done
