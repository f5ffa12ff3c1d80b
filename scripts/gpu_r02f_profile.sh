#!/bin/bash
# r02 final build: launch list of the bench command + ncu --set full of the decode kernels that changed after the
# first r02 captures (overlapped launches, all-CTA combine, staged CUDA-core variant).  Same recipe as
# scripts/gpu_r02_profile.sh; under ncu launches are serialised, so the overlap of consecutive launches is NOT in
# these numbers.
mkdir -p gpurun_out
K='regex:decode_|fmha_|copy4d|paged_|prologue|rope|rms_norm|masked_rows|mask_tile|peer_wait|seqshard|omx_counter|ll_exchange'
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 1500 --csv --log-file gpurun_out/r02f_launches_bench.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu --min-seconds 0 > gpurun_out/r02f_launches_bench.out 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/r02f_launches_bench.csv)"
prof() {  # name, kernel regex, skip, bench args
  ncu --set full --clock-control none -k "regex:$2" -s $3 -c 2 -o gpurun_out/prof_$1 -f python bench.py $4 --steps 4 --warmup 3 --no-cpu --min-seconds 0 > gpurun_out/prof_$1.out 2>&1
  ncu -i gpurun_out/prof_$1.ncu-rep --page raw --csv > gpurun_out/r02f_prof_$1_raw.csv 2>/dev/null
  python - "$1" <<'PY'
import csv, sys
name = sys.argv[1]
rows = list(csv.reader(open(f"gpurun_out/r02f_prof_{name}_raw.csv")))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__cluster_size", "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"]
with open(f"gpurun_out/r02f_prof_{name}_summary.txt", "w") as f:
    for d in data:
        for w in want:
            if w in hdr:
                i = hdr.index(w)
                f.write(f"{w} [{units[i]}] = {d[i][:110]}\n")
        f.write("--\n")
print(open(f"gpurun_out/r02f_prof_{name}_summary.txt").read())
PY
  rm -f gpurun_out/prof_$1.ncu-rep gpurun_out/r02f_prof_$1_raw.csv
}
prof decode_c2 decode_hmma 6 "--workload c2"
prof decode_c2_b8 decode_hmma 6 "--workload c2 --batch 8"
prof decode_c5 decode_hmma 6 "--workload c5"
prof decode_c1 decode_simt 6 "--workload c1"
cuobjdump -sass ominix-mlx_b200/libomx_attn.so | grep -oE "\b(UTCHMMA|UTCBAR|LDTM|STTM|UTMALDG|UTMAPF|UBLKCP|HMMA\.[0-9]+|MOVM|LDSM|SYNCS|MUFU\.EX2|UCGABAR|ACQBULK|ELECT|UTCATOMSWS)\b[A-Z0-9_.]*" | sort | uniq -c | sort -rn > gpurun_out/r02f_sass_counts.txt
head -30 gpurun_out/r02f_sass_counts.txt
