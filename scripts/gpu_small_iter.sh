#!/bin/bash
# Checkpoint for the single-sequence decode latency work: full GPU suite, microbench (cluster on/off), loop timing,
# C2/C5 regression check.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/g_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/g_pytest.log
tail -6 gpurun_out/g_pytest.log
for cl in 0 1; do echo "CLUSTER=$cl"; OMX_DECODE_CLUSTER=$cl timeout 300 python scripts/bench_small_decode.py 2>&1 | tail -4 | tee gpurun_out/small_decode_cluster$cl.jsonl; done
timeout 600 python scripts/bench_decode_loop.py --steps 200 > gpurun_out/g_loop.jsonl 2> gpurun_out/g_loop.err; tail -3 gpurun_out/g_loop.err
python - <<'PY'
import json
for l in open('gpurun_out/g_loop.jsonl'):
    d=json.loads(l)
    print(d['shape'], 'eager us/layer %.2f graph us/layer %.2f  graph GB/s %.0f same=%s' % (d['eager']['us_per_layer'], d['graph']['us_per_layer'], d['graph_hbm_gbs'], d['same_result']))
PY
for wl in c2 c5 c1; do
  echo "$wl $(timeout 300 python bench.py --workload $wl --steps 200 --warmup 10 --no-cpu --graph --rotate 16 2>/dev/null | tail -1 | grep -o '"ms_per_step": [0-9.e-]*' | head -1)"
done
