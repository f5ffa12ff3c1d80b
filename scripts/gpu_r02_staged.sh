#!/bin/bash
out=gpurun_out/r02_staged.log
: > $out
python -m pytest tests/test_decode_gpu.py tests/test_decode_random_gpu.py tests/test_graph_decode_gpu.py tests/test_sdpa_gpu.py tests/test_parallel_gpu.py -x -q 2>&1 | tail -2 | tee -a $out
for st in 1 0 1 0; do
  echo "== OMX_DECODE_STAGED=$st" | tee -a $out
  OMX_BENCH_LABELS=fused,fused_norm OMX_DECODE_STAGED=$st timeout 300 python scripts/bench_small_decode.py "c1 fp32" 2>&1 | grep shape | tee -a $out
  r=$(OMX_DECODE_STAGED=$st timeout 120 python bench.py --workload c1 --steps 20 --warmup 5 --no-cpu --min-seconds 0.25 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step']*1e3,2), round(d['ms_per_step_min']*1e3,2))")
  echo "c1 bench us(median,min): $r" | tee -a $out
done
OMX_BENCH_LABELS=fused OMX_DECODE_TRACE=1 timeout 200 python scripts/bench_small_decode.py "c1 fp32" 2>&1 | grep "decode trace" | head -40 | tail -2 | tee -a $out
