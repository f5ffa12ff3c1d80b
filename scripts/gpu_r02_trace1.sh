#!/bin/bash
out=gpurun_out/r02_trace1.log
: > $out
OMX_BENCH_LABELS=fused timeout 300 python scripts/bench_small_decode.py 2>&1 | grep shape | tee -a $out
for sh in "one rank" "c1 fp32" "8b bf16"; do
OMX_BENCH_LABELS=fused OMX_DECODE_TRACE=1 timeout 200 python scripts/bench_small_decode.py "$sh" 2>&1 | grep "decode trace" | head -40 | tail -2 | tee -a $out
done
