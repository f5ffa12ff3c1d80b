#!/bin/bash
# C1 / C5 latency sweeps over the split-K plan (1 GPU). Output: gpurun_out/s5_small.log
mkdir -p gpurun_out
L=gpurun_out/s5_small.log; : > $L
for wl in c1 c5; do
  for sp in 0 4 8 16 24 32 48 64; do
    for rot in 1 16; do
      r=$(OMX_DECODE_SPLITS=$sp python bench.py --workload $wl --graph --rotate $rot --steps 3200 --no-cpu 2>/dev/null | grep -o '"ms_per_step": [0-9.e-]*' | head -1)
      echo "$wl splits=$sp rotate=$rot $r" >> $L
    done
  done
done
cat $L
