#!/bin/bash
# r02: warm-up walk of the all-CTA combine's code by the idle producer warp, on / off, one box
out=gpurun_out/r02_warm_ab.log
: > $out
python -m pytest tests/test_decode_gpu.py tests/test_decode_random_gpu.py tests/test_graph_decode_gpu.py tests/test_parallel_gpu.py tests/test_paged_gpu.py -x -q 2>&1 | tail -2 | tee -a $out
for g in 1 0 1 0; do
  echo "== OMX_DECODE_WARM=$g" | tee -a $out
  OMX_BENCH_LABELS=fused OMX_DECODE_WARM=$g timeout 300 python scripts/bench_small_decode.py 2>&1 | grep shape | grep -v "c1 fp32" | tee -a $out
  for w in c5; do
    r=$(OMX_DECODE_WARM=$g timeout 120 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu --min-seconds 0.25 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step']*1e3,2), round(d['ms_per_step_min']*1e3,2))")
    echo "$w warm=$g us(median,min): $r" | tee -a $out
  done
done
OMX_BENCH_LABELS=fused OMX_DECODE_TRACE=1 timeout 200 python scripts/bench_small_decode.py "one rank" 2>&1 | grep "decode trace" | head -40 | tail -2 | tee -a $out
