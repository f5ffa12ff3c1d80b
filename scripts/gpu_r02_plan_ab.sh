#!/bin/bash
# r02: plan choices for single-sequence shapes after the code-size work, one box
out=gpurun_out/r02_plan_ab.log
: > $out
for cfg in "OMX_X=1" "OMX_DECODE_GSYNC=2" "OMX_DECODE_CLUSTER=0" "OMX_DECODE_SPLITS=16" "OMX_DECODE_SPLITS=17" "OMX_X=1"; do
  echo "== $cfg" | tee -a $out
  env $cfg OMX_BENCH_LABELS=fused,fused_norm timeout 300 python scripts/bench_small_decode.py 2>&1 | grep shape | tee -a $out
done
