#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parallel_gpu.py -x -q > gpurun_out/s2_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/s2_pytest.log
tail -12 gpurun_out/s2_pytest.log
N=$(nvidia-smi -L | wc -l)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 scripts/bench_seqshard.py 2>gpurun_out/s2_bench.err | tee gpurun_out/seqshard_n$N.json
tail -5 gpurun_out/s2_bench.err
