#!/bin/bash
# final-build bench line at N GPUs, launched the way the driver does it
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
  bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02f_bench_n$N.json 2> gpurun_out/r02f_bench_n$N.err
tail -c 400 gpurun_out/r02f_bench_n$N.err
python scripts/show_bench.py gpurun_out/r02f_bench_n$N.json 2>&1 | tail -14
