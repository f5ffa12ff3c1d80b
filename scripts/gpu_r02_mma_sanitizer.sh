#!/bin/bash
# compute-sanitizer over the mma.sync attention kernel's tests (cp.async stages, named-barrier score exchange, key-group
# hand-over through the stages, split + merge launches, graph-mode position reads) and the prologue kernel's new paths
SUB="tests/test_sdpa_gpu.py::test_mma_all_masks tests/test_sdpa_gpu.py::test_mma_mla_decode_key_groups tests/test_sdpa_gpu.py::test_mma_split_keys_long_context_and_masked_rows tests/test_sdpa_gpu.py::test_mma_strided_views_and_refusals tests/test_sdpa_gpu.py::test_absorbed_mla_glm47_flash_dk576_dv512 tests/test_sdpa_gpu.py::test_mma_decode_key_groups_every_width tests/test_norm_gpu.py::test_fused_decode_with_q_k_norm tests/test_graph_decode_gpu.py::test_dynamic_position_equals_host_offset_step tests/test_prologue_gpu.py"
for tool in memcheck racecheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool --target-processes all --error-exitcode 86 --log-file gpurun_out/sanitizer3_$tool.log python -m pytest $SUB -x -q -p no:cacheprovider > gpurun_out/sanitizer3_${tool}_pytest.log 2>&1
  echo "$tool rc=$? $(tail -1 gpurun_out/sanitizer3_${tool}_pytest.log) | $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer3_$tool.log)"
done
grep -E "hazard|Invalid|error" gpurun_out/sanitizer3_racecheck.log | sort | uniq -c | sort -rn | head -8
