#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parallel_gpu.py tests/test_decode_gpu.py tests/test_graph_decode_gpu.py -x -q > gpurun_out/s_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/s_pytest.log
tail -25 gpurun_out/s_pytest.log
