#!/bin/bash
# r02 A/B (2): pre-wait L2 request depth (l2_early) per grid type, one box
out=gpurun_out/r02_l2_ab2.log
: > $out
python -m pytest tests/test_decode_gpu.py -x -q 2>&1 | tail -3 | tee -a $out
run() {  # workload B early
  r=$(OMX_DECODE_L2EARLY=$3 timeout 120 python bench.py --workload $1 --batch $2 --steps 20 --warmup 5 --no-cpu --min-seconds 0.25 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step']*1e3,2), round(d['ms_per_step_min']*1e3,2), round(d['roofline']['achieved']))")
  echo "$1 B=$2 early=$3 us(median,min),GB/s: $r" | tee -a $out
}
for B in 8 16; do for e in 0 4 6 8 10 12 0 8; do run c2 $B $e; done; done
for B in 32 64; do for e in 0 2 4 6 0; do run c2 $B $e; done; done
for e in 0 4 8 0 8; do run c5 1 $e; done
