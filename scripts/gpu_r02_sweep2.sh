#!/bin/bash
out=gpurun_out/r02_sweep2.log
: > $out
for B in 8 16; do
  for cfg in "1 0" "0 2" "0 3" "0 4" "1 2"; do
    set -- $cfg
    if [ $2 = 0 ]; then unset OMX_DECODE_SPLITS; else export OMX_DECODE_SPLITS=$2; fi
    r=$(OMX_DECODE_CFG=$1 timeout 120 python bench.py --workload c2 --batch $B --steps 20 --warmup 5 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step']*1e3,2), round(d['ms_per_step_min']*1e3,2), round(d['roofline']['achieved']))")
    echo "B=$B cfg=$1 splits=$2 us(median,min),GB/s: $r" | tee -a $out
  done
done
