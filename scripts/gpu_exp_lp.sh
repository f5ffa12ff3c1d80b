#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/exp_lp.log; : > $L
for lib in libomx_attn.so libomx_attn_lp.so; do
  echo "=== $lib" >> $L
  OMX_ATTN_LIB=$PWD/ominix-mlx_b200/$lib python scripts/exp_fmha_err.py >> $L 2>&1
  for wl in c3 c4; do
    OMX_ATTN_LIB=$PWD/ominix-mlx_b200/$lib python bench.py --workload $wl --steps 30 --warmup 5 --no-cpu 2>>$L | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$wl', d['ms_per_step'], d['value'])" >> $L
  done
done
cat $L
