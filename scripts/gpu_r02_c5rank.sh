#!/bin/bash
# r02: one rank of the 8-way kv-head-sharded C5 (4 q / 1 kv head, ctx 32768): graph-replayed latency and per-CTA timeline
out=gpurun_out/r02_c5rank.log
: > $out
for s in 0 32 48 64; do
  echo "== OMX_DECODE_SPLITS=$s" | tee -a $out
  OMX_DECODE_SPLITS=$s timeout 200 python scripts/bench_small_decode.py "one rank" 2>&1 | tail -1 | tee -a $out
done
OMX_BENCH_LABELS=fused OMX_DECODE_TRACE=1 timeout 200 python scripts/bench_small_decode.py "one rank" 2>&1 | grep "decode trace" | head -40 | tail -3 | tee -a $out
