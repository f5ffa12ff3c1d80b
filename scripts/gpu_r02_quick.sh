#!/bin/bash
# r02 quick check: decode tests + C2 shape at 8/16/32/64 rows per GPU + c1/c5 (graph replay, >= 0.5 s timed)
out=gpurun_out/r02_quick.log
: > $out
python -m pytest tests/test_decode_gpu.py tests/test_paged_gpu.py tests/test_graph_decode_gpu.py tests/test_decode_random_gpu.py tests/test_sdpa_gpu.py tests/test_parallel_gpu.py -x -q 2>&1 | tail -3 | tee -a $out
for B in 8 16 32 64; do
  r=$(timeout 120 python bench.py --workload c2 --batch $B --steps 20 --warmup 5 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step']*1e3,2), round(d['ms_per_step_min']*1e3,2), round(d['roofline']['achieved']), round(d['e2e']['ms_per_step']*1e3,2))")
  echo "B=$B us(median,min),GB/s,e2e_us: $r" | tee -a $out
done
for w in c1 c5; do
  r=$(timeout 120 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step']*1e3,2), round(d['ms_per_step_min']*1e3,2))")
  echo "$w us(median,min): $r" | tee -a $out
done
OMX_DECODE_TRACE=1 timeout 120 python bench.py --workload c2 --batch 8 --steps 2 --warmup 3 --no-cpu --eager-e2e 2>&1 | grep "omx decode trace" | tail -2 | tee -a $out
