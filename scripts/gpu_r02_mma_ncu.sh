#!/bin/bash
for w in decode prefill; do
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:sdpa_mma_kernel -s 2 -c 1 -f -o gpurun_out/r02_mma_$w python scripts/gpu_r02_mma_ncu.py $w > gpurun_out/r02_mma_ncu_$w.log 2>&1
  tail -2 gpurun_out/r02_mma_ncu_$w.log
done
ls -la gpurun_out/*.ncu-rep
