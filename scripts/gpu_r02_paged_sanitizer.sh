#!/bin/bash
# compute-sanitizer over the paged-cache tests (incl. the prologue + mma.sync route's page-table reads and length updates)
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --target-processes all --error-exitcode 86 --log-file gpurun_out/sanitizer4_$tool.log python -m pytest tests/test_paged_gpu.py -x -q -p no:cacheprovider > gpurun_out/sanitizer4_${tool}_pytest.log 2>&1
  echo "$tool rc=$? $(tail -1 gpurun_out/sanitizer4_${tool}_pytest.log) | $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer4_$tool.log)"
done
