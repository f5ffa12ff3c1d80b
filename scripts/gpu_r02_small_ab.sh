#!/bin/bash
# r02: single-sequence shapes, all-CTA combine on / off, one box; then the C2 split sizes (regression check)
out=gpurun_out/r02_small_ab.log
: > $out
python -m pytest tests/test_decode_gpu.py tests/test_decode_random_gpu.py tests/test_parallel_gpu.py tests/test_graph_decode_gpu.py tests/test_paged_gpu.py tests/test_sdpa_gpu.py tests/test_norm_gpu.py -x -q 2>&1 | tail -2 | tee -a $out
for g in 1 0 1 0; do
  echo "== OMX_DECODE_GSYNC=$g" | tee -a $out
  OMX_BENCH_LABELS=fused,fused_norm OMX_DECODE_GSYNC=$g timeout 300 python scripts/bench_small_decode.py 2>&1 | grep shape | tee -a $out
done
for B in 8 64; do
  r=$(timeout 120 python bench.py --workload c2 --batch $B --steps 20 --warmup 5 --no-cpu --min-seconds 0.25 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step']*1e3,2), round(d['ms_per_step_min']*1e3,2), round(d['roofline']['achieved']))")
  echo "c2 B=$B us(median,min),GB/s: $r" | tee -a $out
done
for w in c1 c5; do
  r=$(timeout 120 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu --min-seconds 0.25 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step']*1e3,2), round(d['ms_per_step_min']*1e3,2))")
  echo "$w us(median,min): $r" | tee -a $out
done
