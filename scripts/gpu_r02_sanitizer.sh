#!/bin/bash
# r02: compute-sanitizer over a reduced -m gpu subset -- the kernels with hand-written synchronisation:
# decode split-K (L2 ticket combine + cluster/DSMEM combine), the TMA/mbarrier rings, paged decode, the tcgen05
# FMHA ring (incl. array masks), peer flags at world 1, graph-mode decode; (later in the round) the all-CTA combine,
# the data + flag exchange at world 1, overlapped launches with early reads, the staged CUDA-core variant (C1).  Logs -> gpurun_out/sanitizer_*.log
mkdir -p gpurun_out
SUBSET="tests/test_paged_gpu.py tests/test_decode_gpu.py::test_group_sizes tests/test_decode_gpu.py::test_ragged_lengths_bf16 \
tests/test_decode_gpu.py::test_config_c1_qwen3_0p6b_fp32_decode tests/test_decode_gpu.py::test_first_token_and_growth_across_step_boundary \
tests/test_sdpa_gpu.py::test_decode_mask_long_context_split_k tests/test_sdpa_gpu.py::test_decode_kernel_families \
tests/test_fmha_gpu.py::test_two_q_tiles_two_kv_tiles tests/test_fmha_gpu.py::test_ragged_lengths tests/test_fmha_gpu.py::test_long_kv_many_ring_wraps \
tests/test_fmha_gpu.py::test_bool_mask_broadcast_shapes_and_random_pattern tests/test_fmha_gpu.py::test_head_dim_64 \
tests/test_dit_gpu.py::test_zimage_additive_mask tests/test_dit_gpu.py::test_fused_dit_block_f32_out_runs_on_tcgen05 \
tests/test_parallel_gpu.py::test_peer_store_path_world1_raw_abi tests/test_parallel_gpu.py::test_seq_sharded_virtual_ranks_one_gpu \
tests/test_parallel_gpu.py::test_ll_exchange_world1_raw_abi tests/test_decode_gpu.py::test_back_to_back_steps_overlapped_launches \
tests/test_decode_gpu.py::test_config_c5_shape_single_sequence_bf16 tests/test_decode_gpu.py::test_config_c2_shape_reduced_batch_bf16 \
tests/test_graph_decode_gpu.py::test_dynamic_position_equals_host_offset_step tests/test_prologue_gpu.py tests/test_kvcache_gpu.py tests/test_rope_gpu.py tests/test_norm_gpu.py"
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool"
  ( time timeout 1500 compute-sanitizer --tool $tool --target-processes all --error-exitcode 86 \
      --log-file gpurun_out/sanitizer_$tool.log python -m pytest $SUBSET -x -q -p no:cacheprovider ) > gpurun_out/sanitizer_${tool}_pytest.log 2>&1
  echo "rc=$?" >> gpurun_out/sanitizer_${tool}_pytest.log
  tail -4 gpurun_out/sanitizer_${tool}_pytest.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|error" gpurun_out/sanitizer_$tool.log | sort | uniq -c | head -12
done
