#!/bin/bash
# Composite prefill / DiT measurements (prologue + attention) and the prologue kernel's ncu numbers.
mkdir -p gpurun_out
L=gpurun_out/s6_composite.log; : > $L
for wl in c3 c4; do
  for comp in "" "--composite"; do
    echo "== $wl $comp" >> $L
    python bench.py --workload $wl $comp --steps 30 --warmup 5 --no-cpu 2>>$L | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['config']['kernel'], d['gpu_launches'], d['clocks'])" >> $L
  done
done
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_c3_composite.csv python bench.py --workload c3 --composite --steps 3 --warmup 3 --no-cpu > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_c4_composite.csv python bench.py --workload c4 --composite --steps 3 --warmup 3 --no-cpu > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:qkv_prologue -s 3 -c 1 -f -o gpurun_out/prof_prologue_c3 python bench.py --workload c3 --composite --steps 3 --warmup 3 --no-cpu > /dev/null 2>&1
cat $L
grep -c . gpurun_out/launches_c3_composite.csv
