#!/bin/bash
# r02 (2 GPUs): parity of the sharded layouts incl. the data + flag exchange, then C5 at N = 2: LL words vs flags vs NCCL
out=gpurun_out/r02_ll_n2.log
: > $out
timeout 600 python -m pytest tests/test_parallel_gpu.py -x -q 2>&1 | tail -4 | tee -a $out
for g in peer peer_flags; do
  OMX_BENCH_C5_GATHER=$g timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --workload c5 --steps 20 --warmup 5 --no-cpu --min-seconds 0.25 2> gpurun_out/r02_ll_n2_$g.err | tail -1 > gpurun_out/r02_ll_n2_$g.json
  python - <<PY | tee -a $out
import json
d=json.load(open("gpurun_out/r02_ll_n2_$g.json"))
print("c5 N=2 gather=$g us/step", round(d["ms_per_step"]*1e3,2), "min", round(d["ms_per_step_min"]*1e3,2), "e2e", round(d["e2e"]["ms_per_step"]*1e3,2), "parity", d.get("parity_check"), "launches/step", d["run"]["launches_per_step"])
PY
done
