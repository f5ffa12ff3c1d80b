#!/bin/bash
# r02: A/B of decode-kernel BUILDS in one box (make variant ...; OMX_ATTN_LIB selects the library)
out=gpurun_out/r02_variants.log
: > $out
for rep in 1 2; do
for v in "" _b _m _c; do
  lib=$PWD/ominix-mlx_b200/libomx_attn$v.so
  [ -f $lib ] || continue
  echo "== lib$v (rep $rep)" | tee -a $out
  OMX_ATTN_LIB=$lib OMX_BENCH_LABELS=fused,fused_norm timeout 300 python scripts/bench_small_decode.py 2>&1 | grep shape | tee -a $out
done
done
