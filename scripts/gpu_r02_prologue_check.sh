#!/bin/bash
timeout 600 python -m pytest tests/test_prologue_gpu.py tests/test_prefill_gpu.py tests/test_dit_gpu.py -x -q 2>&1 | tail -3
python scripts/gpu_r02_prologue_ncu.py 10 2>&1 | tail -3
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:qkv_prologue -c 3 --csv --log-file gpurun_out/r02_prologue_launches.csv python scripts/gpu_r02_prologue_ncu.py 1 > /dev/null 2>&1
grep qkv_prologue gpurun_out/r02_prologue_launches.csv | awk -F'","' '{print $(NF-3), $(NF-2), $NF}' | sed 's/"//g'
