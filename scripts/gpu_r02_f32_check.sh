#!/bin/bash
out=gpurun_out/r02_f32.log
: > $out
python -m pytest tests/test_sdpa_gpu.py tests/test_decode_gpu.py tests/test_paged_gpu.py tests/test_dit_gpu.py tests/test_prefill_gpu.py -x -q 2>&1 | tail -3 | tee -a $out
SUB="tests/test_paged_gpu.py::test_c2_geometry_with_norms_bf16 tests/test_decode_gpu.py::test_config_c2_shape_reduced_batch_bf16 tests/test_sdpa_gpu.py::test_f32_tiled_all_masks tests/test_sdpa_gpu.py::test_f32_tiled_strided_views_and_fully_masked_rows tests/test_parallel_gpu.py::test_ll_exchange_world1_raw_abi tests/test_decode_gpu.py::test_back_to_back_steps_overlapped_launches"
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --target-processes all --error-exitcode 86 --log-file gpurun_out/sanitizer2_$tool.log python -m pytest $SUB -x -q -p no:cacheprovider > gpurun_out/sanitizer2_${tool}_pytest.log 2>&1
  echo "$tool rc=$? $(tail -1 gpurun_out/sanitizer2_${tool}_pytest.log) | $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer2_$tool.log)" | tee -a $out
done
python scripts/gpu_r02_f32_prefill.py 2>&1 | tee -a $out
ominix-mlx_b200/host/test_host --bench 2>&1 | tee -a $out
