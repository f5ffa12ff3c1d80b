#!/bin/bash
# in-box A/B of builds of the library on the sdpa_mma timing script (OMX_ATTN_LIB selects the .so)
for lib in ${LIBS:-libomx_attn_old.so libomx_attn.so}; do
  echo "== $lib"
  OMX_ATTN_LIB=$PWD/ominix-mlx_b200/$lib timeout 200 python scripts/gpu_r02_mma.py 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print('  %-70s %8.4f ms %7.1f TFLOP/s' % (d['shape'][:70], d['sdpa_mma']['ms'], d['sdpa_mma']['TFLOP/s']))"
done
