#!/bin/bash
# r02 A/B: L2 prefetch distance of the TMA decode kernel (in-loop l2_ahead, pre-wait l2_early), one box
out=gpurun_out/r02_l2_ab.log
: > $out
run() {  # B ahead early
  r=$(OMX_DECODE_L2AHEAD=$2 OMX_DECODE_L2EARLY=$3 timeout 120 python bench.py --workload c2 --batch $1 --steps 20 --warmup 5 --no-cpu --min-seconds 0.25 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step']*1e3,2), round(d['ms_per_step_min']*1e3,2), round(d['roofline']['achieved']))")
  echo "B=$1 ahead=$2 early=$3 us(median,min),GB/s: $r" | tee -a $out
}
for B in 8 16; do
  for cfg in "0 0" "2 0" "4 0" "8 0" "16 0" "0 8" "0 16" "4 8" "8 16" "0 0"; do
    set -- $cfg
    run $B $1 $2
  done
done
for cfg in "0 0" "4 0" "8 0" "0 16" "8 16" "0 0"; do
  set -- $cfg
  run 64 $1 $2
  run 32 $1 $2
done
