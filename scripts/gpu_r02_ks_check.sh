#!/bin/bash
timeout 900 python -m pytest tests/test_sdpa_gpu.py tests/test_decode_gpu.py tests/test_norm_gpu.py tests/test_decode_random_gpu.py -x -q 2>&1 | tail -4
timeout 300 python scripts/gpu_r02_simt_vs_mma.py 2>&1 | cut -c1-330
timeout 300 python scripts/gpu_r02_d256_decode.py 2>&1 | cut -c1-700
