#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parallel_gpu.py -x -q 2>&1 | tail -4
N=$(nvidia-smi -L | wc -l)
for lay in head seq; do
  for gr in "" "--graph"; do
    echo "== c5 N=$N layout=$lay $gr: $(timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --workload c5 --layout $lay --gather peer $gr --steps 320 --warmup 10 --no-cpu 2>gpurun_out/gs_err.log | tail -1 | grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.e-]*' | head -2 | tr '\n' ' ')"
    tail -2 gpurun_out/gs_err.log | grep -i "error\|Traceback" | head -2
  done
done
