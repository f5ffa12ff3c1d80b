#!/bin/bash
# Decode-path checkpoint: decode / graph / golden / norm tests, loop timing, C2/C5/C1 regression check.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_decode_gpu.py tests/test_graph_decode_gpu.py tests/test_golden_gpu.py tests/test_norm_gpu.py tests/test_sdpa_gpu.py tests/test_parallel_gpu.py -x -q > gpurun_out/g_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/g_pytest.log
tail -30 gpurun_out/g_pytest.log
timeout 600 python scripts/bench_decode_loop.py --steps 200 > gpurun_out/g_loop.jsonl 2> gpurun_out/g_loop.err; tail -3 gpurun_out/g_loop.err
python - <<'PY'
import json
for l in open('gpurun_out/g_loop.jsonl'):
    d=json.loads(l)
    print(d['shape'], 'eager us/layer %.2f graph us/layer %.2f  graph GB/s %.0f same=%s' % (d['eager']['us_per_layer'], d['graph']['us_per_layer'], d['graph_hbm_gbs'], d['same_result']))
PY
for wl in c2 c5 c1; do
  for kpw in 0 1; do
  echo "$wl KPW=$kpw $(OMX_DECODE_KPW=$kpw timeout 300 python bench.py --workload $wl --steps 400 --warmup 10 --no-cpu --graph --rotate 16 2>/dev/null | tail -1 | grep -o '"ms_per_step": [0-9.e-]*')"
  [ $wl != c1 ] && break
  done
done
