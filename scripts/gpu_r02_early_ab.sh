#!/bin/bash
# r02 A/B: programmatic dependent launch with the next launch's TMA ring filled BEFORE its dependency wait
# (stable_rows): C2 shape at the per-GPU batches of the split, c5, c1.  One box, interleaved.
out=gpurun_out/r02_early_ab.log
: > $out
python -m pytest tests/test_decode_gpu.py tests/test_paged_gpu.py tests/test_graph_decode_gpu.py tests/test_decode_random_gpu.py tests/test_sdpa_gpu.py tests/test_kvcache_gpu.py tests/test_parallel_gpu.py -x -q 2>&1 | tail -3 | tee -a $out
for B in 8 16 32 64; do
  for cfg in "0 0" "1 0" "1 1" "0 0" "1 1"; do
    set -- $cfg
    r=$(OMX_DECODE_PDL=$1 OMX_DECODE_EARLY=$2 timeout 120 python bench.py --workload c2 --batch $B --steps 20 --warmup 5 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step']*1e3,2), round(d['ms_per_step_min']*1e3,2), round(d['roofline']['achieved']), round(d['e2e']['ms_per_step']*1e3,2))")
    echo "B=$B pdl=$1 early=$2 us(median,min),GB/s,e2e_us: $r" | tee -a $out
  done
done
for w in c5 c1 c2_paged; do
  for cfg in "0 0" "1 0" "1 1"; do
    set -- $cfg
    r=$(OMX_DECODE_PDL=$1 OMX_DECODE_EARLY=$2 timeout 120 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step']*1e3,2), round(d['ms_per_step_min']*1e3,2))")
    echo "$w pdl=$1 early=$2 us(median,min): $r" | tee -a $out
  done
done
