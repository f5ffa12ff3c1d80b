#!/bin/bash
# r02 A/B of two builds of the library on the prefill / DiT workloads, interleaved in ONE box (power state and
# box-to-box variance are larger than the effects): OMX_ATTN_LIB selects the build.
out=gpurun_out/r02_fmha_ab.log
: > $out
python -m pytest tests/test_fmha_gpu.py tests/test_dit_gpu.py tests/test_prefill_gpu.py tests/test_prologue_gpu.py -x -q 2>&1 | tail -3 | tee -a $out
NEW=$PWD/ominix-mlx_b200/libomx_attn.so
OLD=$PWD/ominix-mlx_b200/libomx_attn_prev.so
for rep in 1 2; do
  for w in c3 c4; do
    for lib in OLD NEW; do
      r=$(OMX_ATTN_LIB=${!lib} timeout 200 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), round(d['ms_per_step_min'],4), round(d['value'],1), round(d['roofline']['achieved_best_block'],1), d['clocks']['sm_mhz'])")
      echo "$w $lib rep$rep ms(median,min) TF(median,best) clk: $r" | tee -a $out
    done
  done
done
