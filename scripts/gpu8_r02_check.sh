#!/bin/bash
# r02 (8 GPUs): the batch split of C2 and the kv-head-sharded C5 (data + flag exchange vs flags) at N = 8
out=gpurun_out/r02_n8_check.log
: > $out
run() {  # tag, env, workload
  env $2 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 8 --workload $3 --steps 20 --warmup 5 --no-cpu 2> gpurun_out/r02_n8_$1.err | tail -1 > gpurun_out/r02_n8_$1.json
  python - <<PY | tee -a $out
import json
d=json.load(open("gpurun_out/r02_n8_$1.json"))
print("$1", "us/step", round(d["ms_per_step"]*1e3,2), "min", round(d.get("ms_per_step_min",0)*1e3,2), "value", round(d["value"]), "e2e us", round(d["e2e"]["ms_per_step"]*1e3,2), "frac", round(d["roofline"]["frac"],3), "parity", d.get("parity_check"), "clk", d["clocks"].get("sm_mhz"), d["clocks"].get("reasons"))
PY
}
run c2 OMX_X=1 c2
run c5_ll OMX_X=1 c5
run c5_flags OMX_BENCH_C5_GATHER=peer_flags c5
run c5_nccl OMX_X=1 c5_collective
