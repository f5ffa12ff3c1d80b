#!/bin/bash
# r02: all-CTA combine with tagged words (no fence, no meeting point) vs the counter variant, one box
out=gpurun_out/r02_gsll_ab.log
: > $out
python -m pytest tests/test_decode_gpu.py tests/test_decode_random_gpu.py tests/test_graph_decode_gpu.py tests/test_sdpa_gpu.py tests/test_parallel_gpu.py tests/test_paged_gpu.py tests/test_host_cpp.py -x -q 2>&1 | tail -2 | tee -a $out
for g in 1 0 1 0; do
  echo "== OMX_DECODE_GSLL=$g" | tee -a $out
  OMX_BENCH_LABELS=fused OMX_DECODE_GSLL=$g timeout 300 python scripts/bench_small_decode.py 2>&1 | grep shape | tee -a $out
  for w in c1 c5; do
    r=$(OMX_DECODE_GSLL=$g timeout 120 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu --min-seconds 0.25 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step']*1e3,2), round(d['ms_per_step_min']*1e3,2))")
    echo "$w gsll=$g us(median,min): $r" | tee -a $out
  done
done
OMX_BENCH_LABELS=fused OMX_DECODE_TRACE=1 timeout 200 python scripts/bench_small_decode.py "one rank" 2>&1 | grep "decode trace" | head -40 | tail -2 | tee -a $out
