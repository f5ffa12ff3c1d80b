#!/bin/bash
# full validation of the build: every GPU test, smoke(), the driver's bench command, refreshed launch list of sdpa_mma
out=gpurun_out/r02g
timeout 900 python -m pytest tests -x -q -m gpu > ${out}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a ${out}_pytest.log
tail -3 ${out}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > ${out}_bench_n1.json 2> ${out}_bench_n1.err; echo "bench rc=$?"
python scripts/show_bench.py ${out}_bench_n1.json 2>&1 | tail -16
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_mma_launches.csv python scripts/gpu_r02_mma_probe.py > /dev/null 2>&1
grep -E "sdpa_mma_kernel" gpurun_out/r02_mma_launches.csv | awk -F'","' '{print $5, $NF}' | sed 's/"//g' | tail -12
