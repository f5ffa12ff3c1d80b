#!/bin/bash
# ncu --set full on the single-sequence decode launches of the loop script (eager mode, cold KV per layer).
mkdir -p gpurun_out
for sh in "0.6b fp32" "8b bf16 B1"; do
  tag=$(echo "$sh" | tr ' .' '__')
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:decode -s 120 -c 3 -f \
    -o gpurun_out/small_$tag python scripts/bench_decode_loop.py --only "$sh" --modes eager --steps 3 --warmup 2 \
    > gpurun_out/small_$tag.log 2>&1
  tail -2 gpurun_out/small_$tag.log
done
ls -la gpurun_out/*.ncu-rep
