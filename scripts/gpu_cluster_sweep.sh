#!/bin/bash
# Cluster-size sweep for the single-sequence decode shapes.
OMX_DECODE_TRACE=1 timeout 200 python scripts/bench_decode_loop.py --only "0.6b fp32" --modes eager --steps 1 --warmup 1 2>&1 | grep "co-resident" | sort | uniq
OMX_DECODE_TRACE=1 timeout 200 python scripts/bench_decode_loop.py --only "8b bf16 B1" --modes eager --steps 1 --warmup 1 2>&1 | grep "co-resident" | sort | uniq
for cap in 0 8 10 12 14 16; do
  echo "CLUSTER_MAX=$cap"
  if [ $cap = 0 ]; then export OMX_DECODE_CLUSTER=0; else export OMX_DECODE_CLUSTER=1 OMX_DECODE_CLUSTER_MAX=$cap; fi
  OMX_DECODE_KPW=1 timeout 300 python scripts/bench_small_decode.py 2>&1 | tail -4 | cut -c1-200
done
