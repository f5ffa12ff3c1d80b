#!/bin/bash
# Round checkpoint on one B200: GPU tests, smoke, the default bench line, the other workloads, composite runs.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/v_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/v_pytest.log
tail -5 gpurun_out/v_pytest.log
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/v_bench_c2.json 2> gpurun_out/v_bench_c2.err; cat gpurun_out/v_bench_c2.json
L=gpurun_out/v_workloads.log; : > $L
for wl in c1 c3 c4 c5; do
  for comp in "" "--composite"; do
    if [ -n "$comp" ] && [ $wl != c3 ] && [ $wl != c4 ]; then continue; fi
    echo "== $wl $comp" >> $L
    timeout 300 python bench.py --workload $wl $comp --steps 30 --warmup 5 --no-cpu 2>>$L | tail -1 >> $L
  done
done
cat $L | cut -c1-600
