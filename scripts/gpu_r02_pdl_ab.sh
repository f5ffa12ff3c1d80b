#!/bin/bash
# r02 A/B: programmatic dependent launch and the push combine, C2 shape at the per-GPU batches of the split
out=gpurun_out/r02_pdl_ab.log
: > $out
python -m pytest tests/test_decode_gpu.py tests/test_paged_gpu.py tests/test_graph_decode_gpu.py tests/test_decode_random_gpu.py tests/test_sdpa_gpu.py -x -q 2>&1 | tail -3 | tee -a $out
for B in 8 16 32 64; do
  for cfg in "1 1" "0 1" "1 0" "0 0"; do
    set -- $cfg
    r=$(OMX_DECODE_PDL=$1 OMX_DECODE_PUSH=$2 timeout 120 python bench.py --workload c2 --batch $B --steps 20 --warmup 5 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step']*1e3,2), round(d['ms_per_step_min']*1e3,2), round(d['roofline']['achieved']), round(d['e2e']['ms_per_step']*1e3,2))")
    echo "B=$B pdl=$1 push=$2 us(median,min),GB/s,e2e_us: $r" | tee -a $out
  done
done
for w in c1 c5; do
  for pdl in 1 0; do
    r=$(OMX_DECODE_PDL=$pdl timeout 120 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step']*1e3,2), round(d['ms_per_step_min']*1e3,2))")
    echo "$w pdl=$pdl us(median,min): $r" | tee -a $out
  done
done
OMX_DECODE_TRACE=1 timeout 120 python bench.py --workload c2 --batch 8 --steps 2 --warmup 3 --no-cpu --eager-e2e 2>&1 | grep "omx decode trace" | tail -2 | tee -a $out
