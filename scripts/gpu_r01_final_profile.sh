#!/bin/bash
# Round-1 closing evidence for the decode kernel after the operand swap / cluster combine (1 GPU).
# The .ncu-rep files exceed what gpurun copies back, so the raw and source pages are exported on the box.
mkdir -p gpurun_out
python bench.py --steps 500 --warmup 20 > gpurun_out/f_bench_c2.json 2> gpurun_out/f_bench_c2.err; cut -c1-300 gpurun_out/f_bench_c2.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/f_launches_c2.csv python bench.py --steps 5 --warmup 3 --no-cpu > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:decode_hmma -s 3 -c 2 -f -o /tmp/f_prof_decode_c2 python bench.py --steps 5 --warmup 3 --no-cpu > /dev/null 2>&1
ncu -i /tmp/f_prof_decode_c2.ncu-rep --page raw --csv > gpurun_out/f_prof_decode_c2_raw.csv 2>/dev/null
ncu -i /tmp/f_prof_decode_c2.ncu-rep --page details --csv > gpurun_out/f_prof_decode_c2_details.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:decode_hmma -s 150 -c 2 -f -o /tmp/f_prof_decode_8b_b1 python scripts/bench_decode_loop.py --only "8b bf16 B1" --modes eager --steps 3 --warmup 2 > /dev/null 2>&1
ncu -i /tmp/f_prof_decode_8b_b1.ncu-rep --page raw --csv > gpurun_out/f_prof_decode_8b_b1_raw.csv 2>/dev/null
cuobjdump -sass ominix-mlx_b200/libomx_attn.so 2>/dev/null | grep -c "MOVM" > gpurun_out/f_movm_count.txt
ls -la gpurun_out/f_*
