#!/bin/bash
# Decode measurements + ncu evidence (1 GPU). Outputs under gpurun_out/.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x --timeout=400 -k "rope_empty or validation or decode" 2>&1 | tail -3
python bench.py --steps 1000 --warmup 20 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -c 1500 gpurun_out/bench_c2.json
OMX_DECODE_CFG=1 python bench.py --steps 1000 --warmup 20 --no-cpu > gpurun_out/bench_c2_cfg1.json 2>&1; grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_c2_cfg1.json | head -1
for sp in 2 8; do OMX_DECODE_SPLITS=$sp python bench.py --steps 300 --warmup 20 --no-cpu 2>&1 | grep -o '"ms_per_step": [0-9.]*' | head -1; done
python bench.py --workload c1 --graph --steps 3200 --no-cpu > gpurun_out/bench_c1_warm.json 2>&1; grep -o '"ms_per_step": [0-9.e-]*\|"achieved": [0-9.]*' gpurun_out/bench_c1_warm.json | head -2
python bench.py --workload c1 --graph --rotate 16 --steps 3200 --no-cpu > gpurun_out/bench_c1_cold.json 2>&1; grep -o '"ms_per_step": [0-9.e-]*\|"achieved": [0-9.]*' gpurun_out/bench_c1_cold.json | head -2
python bench.py --workload c5 --graph --steps 3200 --no-cpu > gpurun_out/bench_c5_warm.json 2>&1; grep -o '"ms_per_step": [0-9.e-]*\|"achieved": [0-9.]*' gpurun_out/bench_c5_warm.json | head -2
python bench.py --workload c5 --graph --rotate 4 --steps 3200 --no-cpu > gpurun_out/bench_c5_cold.json 2>&1; grep -o '"ms_per_step": [0-9.e-]*\|"achieved": [0-9.]*' gpurun_out/bench_c5_cold.json | head -2
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 5 --warmup 3 --no-cpu > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:decode_hmma -s 3 -c 2 -f -o gpurun_out/prof_decode_c2 python bench.py --steps 5 --warmup 3 --no-cpu > /dev/null 2>&1
ls -la gpurun_out/
