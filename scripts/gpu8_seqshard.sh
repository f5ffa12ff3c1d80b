#!/bin/bash
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 scripts/bench_seqshard.py 2>gpurun_out/s8_bench.err | tee gpurun_out/seqshard_n$N.json
tail -3 gpurun_out/s8_bench.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 200 --warmup 10 2>/dev/null | tail -1 | tee gpurun_out/bench_c2_n$N.json | cut -c1-330
