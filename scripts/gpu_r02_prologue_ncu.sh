#!/bin/bash
python scripts/gpu_r02_prologue_ncu.py 10 2>&1 | tail -3
timeout 300 ncu --set full --import-source on --clock-control none -k regex:qkv_prologue -s 2 -c 1 -f -o gpurun_out/r02_prologue python scripts/gpu_r02_prologue_ncu.py 2 > gpurun_out/r02_prologue_ncu.log 2>&1
tail -2 gpurun_out/r02_prologue_ncu.log
