#!/bin/bash
out=gpurun_out/r02_mma.log
: > $out
timeout 240 python -m pytest tests/test_sdpa_gpu.py tests/test_decode_gpu.py -x -q 2>&1 | tail -15 | tee -a $out
timeout 200 python scripts/gpu_r02_mma.py 2>&1 | tee -a $out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_mma_launches.csv python scripts/gpu_r02_mma_probe.py > /dev/null 2>&1
grep -E "sdpa_mma|combine" gpurun_out/r02_mma_launches.csv | awk -F'","' '{print $5, $NF}' | sed 's/"//g' | tee -a $out
