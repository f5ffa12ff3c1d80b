#!/bin/bash
# r02: how the C2 shape behaves at the per-GPU batches of the strong-scaling split (B = 32 / 16 / 8),
# over forced split counts and the two pipeline variants; graph replay so the host is out of the picture.
mkdir -p gpurun_out
out=gpurun_out/r02_split_sweep.log
: > $out
for B in 8 16 32; do
  for cfg in 0 1; do
    for sp in 0 1 2 3 4 5 6 8; do
      if [ $sp = 0 ]; then unset OMX_DECODE_SPLITS; else export OMX_DECODE_SPLITS=$sp; fi
      r=$(OMX_DECODE_CFG=$cfg timeout 120 python bench.py --batch $B --steps 320 --warmup 5 --no-cpu --graph 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step']*1e3,2), round(d['roofline']['achieved']))")
      echo "B=$B cfg=$cfg splits=$sp us,GB/s: $r" | tee -a $out
    done
  done
done
unset OMX_DECODE_SPLITS
echo "== default plan, eager vs graph" | tee -a $out
for B in 8 16 32 64; do
  for g in "" "--graph"; do
    r=$(timeout 120 python bench.py --batch $B --steps 320 --warmup 5 --no-cpu $g 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step']*1e3,2), round(d['roofline']['achieved']), round(d['e2e']['ms_per_step']*1e3,2))")
    echo "B=$B default $g us,GB/s,e2e_us: $r" | tee -a $out
  done
done
OMX_DECODE_TRACE=1 timeout 120 python bench.py --batch 8 --steps 3 --warmup 3 --no-cpu 2>&1 | grep "omx decode" | tail -3 | tee -a $out
