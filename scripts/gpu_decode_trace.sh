#!/bin/bash
# Per-CTA phase timelines of the single-sequence decode launches (OMX_DECODE_TRACE=1|2).
mkdir -p gpurun_out
for sh in "0.6b fp32" "0.6b bf16" "8b bf16 B1" "mixtral"; do
  echo "== $sh"
  tag=$(echo "$sh" | tr ' .' '__')
  OMX_DECODE_TRACE=2 timeout 300 python scripts/bench_decode_loop.py --only "$sh" --modes eager --steps 1 --warmup 2 2>&1 | grep "decode trace\|co-resident" > gpurun_out/trace_$tag.log
  grep -v "trace cta" gpurun_out/trace_$tag.log | tail -2
done
