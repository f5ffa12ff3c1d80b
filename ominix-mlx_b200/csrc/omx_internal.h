// omx_internal.h -- cross-translation-unit declarations of libomx_attn (not installed).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/omx_attn.h"

namespace omx {

// ---- rope.cu ----
struct RopeTableRef {
  const float* cos = nullptr;  // [n_pos, half] float32, device
  const float* sin = nullptr;
  int half = 0;
  int n_pos = 0;
  // the same rows rounded once to bf16 ([0]) / f16 ([1]), [n_pos, half]
  const void* cos16[2] = {nullptr, nullptr};
  const void* sin16[2] = {nullptr, nullptr};
};
RopeTableRef get_rope_table(int dims, bool has_base, float base, float scale,
                            const float* freqs_host, int need_positions, cudaStream_t stream);
void rope_forward(const omx_array* out, const omx_array* x, int dims, bool traditional,
                  omx_optional_float base, float scale, int offset, const omx_array* offset_arr,
                  int max_position, const omx_array* freqs, cudaStream_t stream);
void dit_rope_forward(const omx_array* out, const omx_array* x, const omx_array* cs,
                      const omx_array* sn, cudaStream_t stream);

// ---- norm.cu ----
void rms_norm_forward(const omx_array* out, const omx_array* x, const omx_array* weight, float eps,
                      cudaStream_t stream);

// ---- prologue.cu ----
// One launch for the row-wise ops in front of the attention kernel: per segment, out = rope(norm(x)).
struct PrologueSeg {
  const omx_array* x = nullptr;  // [B,H,L,D] logical view (any leading strides, feature axis contiguous)
  omx_array out{};               // same shape; may point into the KV cache / a joint buffer
  const omx_array* w = nullptr;  // RMSNorm weight [D] or null
  bool rope = false;             // rotate (true) or copy
  int tok0 = 0;                  // table row of token l == 0
  int64_t dyn_row_stride = 0;    // graph mode: `out` moves by this many elements per position (cache rows)
  bool paged_dst = false;        // paged mode: `out` is the page pool ([page][H][64][D] through strides 0 / 1 / 2)
};
struct PrologueCall {
  PrologueSeg seg[6];
  int nseg = 0;
  int dims = 0;  // rotated features (mode 1); mode 2 rotates all D
  bool traditional = false;
  int mode = 1;  // 1: fast::rope position table; 2: per-token tables in the array dtype (DiT)
  RopeTableRef table;
  const omx_array *tcos = nullptr, *tsin = nullptr;  // mode 2
  int64_t tcs[3] = {0, 0, 0}, tss[3] = {0, 0, 0};    // element strides of (b, token, pair)
  float eps = 0.f;
  // graph mode: *pos_dev (device) is added to every segment's tok0 and, times dyn_row_stride, to its destination
  const int* pos_dev = nullptr;
  // paged mode (single-token steps on the paged cache): sequence b stands at lens_in[b] rows (< 0: released slot, its
  // rows are skipped); that is its rope row, and the page row its k' / v go to (segments with paged_dst) comes from the
  // block table; the k segment's head-0 rows store lens_in[b] + 1 into lens_out[b]
  const struct PagedRef* paged = nullptr;
};
// false: shape / layout outside the kernel's coverage (nothing launched) -- the caller composes the
// standalone ops instead.
bool qkv_prologue(const PrologueCall& c, cudaStream_t stream);

// ---- kv_cache.cu ----
struct KVCacheImpl;
KVCacheImpl* kv_cache_create(int step, bool concat);
void kv_cache_destroy(KVCacheImpl* c);
int kv_cache_offset(const KVCacheImpl* c);
// rows the launch of the call that just updated the cache may read before its dependency wait (decode.cu, PDL)
int kv_cache_stable_rows(const KVCacheImpl* c);
void kv_cache_reset(KVCacheImpl* c);
void kv_cache_reserve(KVCacheImpl* c, int rows);
int kv_cache_trim(KVCacheImpl* c, int n);
bool kv_cache_is_concat(const KVCacheImpl* c);
// Grows if needed (reference rule), copies the n new rows unless `skip_copy` (the fused decode
// kernel writes them itself), advances the offset and fills the fetched views.
void kv_cache_update(KVCacheImpl* c, const omx_array* keys, const omx_array* values,
                     omx_array* keys_out, omx_array* values_out, bool skip_copy,
                     cudaStream_t stream);
void kv_cache_state(const KVCacheImpl* c, omx_array* kbuf, omx_array* vbuf);
// The composite entry points advance the cache (skip_copy) BEFORE the launch that writes the rows; if anything
// between the two throws, the guard puts offset / capacity back so that a failed call leaves the cache as it
// found it (no phantom row that a later step would attend).
struct KVCacheSnapshot {
  int offset;
  int64_t cap;
  bool has;
};
KVCacheSnapshot kv_cache_snapshot(const KVCacheImpl* c);
void kv_cache_rollback(KVCacheImpl* c, const KVCacheSnapshot& s, cudaStream_t stream);
struct KVCacheTxn {
  KVCacheImpl* c;
  KVCacheSnapshot snap;
  cudaStream_t stream;
  bool done = false;
  KVCacheTxn(KVCacheImpl* c_, cudaStream_t s) : c(c_), snap(kv_cache_snapshot(c_)), stream(s) {}
  void commit() { done = true; }
  ~KVCacheTxn() {
    if (!done) kv_cache_rollback(c, snap, stream);
  }
};
// Graph mode: pin the buffers at >= max_rows physical rows (never-written tail zeroed) and allocate the
// cache-owned scratch; the addresses stay fixed until the cache grows past max_rows.
void kv_cache_prepare_graph(KVCacheImpl* c, int max_rows, size_t scratch_bytes, cudaStream_t stream);
// Views over all max_rows pinned rows + the scratch; false if prepare_graph has not been called.
bool kv_cache_graph_view(const KVCacheImpl* c, omx_array* k, omx_array* v, void** scratch,
                         size_t* scratch_bytes, int* max_rows);
// Host bookkeeping for n rows appended by dynamic-position launches (reference growth rule, no copy).
void kv_cache_advance(KVCacheImpl* c, int n, cudaStream_t stream);
void kv_cache_shape(const KVCacheImpl* c, int* B, int* H, int* Dk, int* Dv, int* dtype);
// Strided 4-D copy (dst, src same shape/dtype); used by caches and tests.
void copy4d(const omx_array* dst, const omx_array* src, cudaStream_t stream);

// ---- workspace (omx_api.cu) ----
// Per-(device, stream) scratch that only grows; zero-initialised on (re)allocation.
void* get_workspace(size_t bytes, cudaStream_t stream);
// Scratch of the composite entry points (kept apart from the kernels' own workspace above).
void* get_outer_workspace(size_t bytes, cudaStream_t stream);
// A second, separately zero-initialised region for self-resetting arrival counters.
int* get_counters(size_t count, cudaStream_t stream);
// Pool of {value, launch tag} words for the tagged all-CTA combine (decode.cu, gsync == 2): only ever holds such
// words (zeroed when allocated), *seq = launches completed on it (device counter, never moves).
void* get_tagged_workspace(size_t bytes, cudaStream_t stream, unsigned** seq);
int sm_count();

// ---- sdpa_generic.cu ----
enum MaskMode { MASK_NONE = 0, MASK_CAUSAL = 1, MASK_BOOL = 2, MASK_ADD = 3 };
struct SdpaArgs {
  const omx_array *out, *q, *k, *v;
  float scale;
  int mask_mode;
  const omx_array* mask;  // broadcastable to [B,Hq,Lq,Lk]
  int64_t mask_strides[4];  // element strides after broadcasting (0 on broadcast axes)
  int B, Hq, Hkv, Lq, Lk, D, Dv;
};
void sdpa_generic(const SdpaArgs& a, cudaStream_t stream);
// Rewrites the rows flagged in dead[B][Hq][Lq] (array mask hides every key) with the reference's
// uniform average over all Lk rows of V; launched by the tile-skipping kernels after their own pass.
void masked_rows_fixup(const SdpaArgs& a, const uint8_t* dead, cudaStream_t stream);

// ---- decode.cu ----
struct DecodeFused {  // optional fused rope + append of the new token (L == 1)
  bool enabled = false;
  const omx_array* k_new = nullptr;  // [B,Hkv,1,D] un-roped
  const omx_array* v_new = nullptr;  // [B,Hkv,1,Dv]
  int rope_dims = 0;                 // 0: no rotation
  bool traditional = false;
  RopeTableRef table;
  int position = 0;  // = cache offset before the append
  // head-sharded output: store this rank's heads into every rank's full [B,Hq_total,1,D] buffer
  const omx_peer_group* peers = nullptr;  // out pointers already shifted to this rank's first head
  int head_offset = 0;
  bool peer_wait = false;  // the launch itself waits for every peer's arrival of this step
  // data + flag exchange (omx_attn_decode_fused_sharded_ll): staging words instead of peer stores + counters;
  // out_full = base of the local full-head output the received words are unpacked into
  const omx_ll_group* ll = nullptr;
  void* ll_out_full = nullptr;
  // per-head RMSNorm of q / k_new before the rotation (Qwen3 q_norm / k_norm): [D] weights in the
  // q dtype, contiguous; null = no norm
  const void* q_norm_w = nullptr;
  const void* k_norm_w = nullptr;
  float norm_eps = 0.f;
  // sequence-sharded decode: append = false on ranks that do not own the new token (k_new / v_new unused);
  // partial = true: `out` is a float32 slot [B][Hq][D + 2] (normalised output of the local keys, m, l)
  bool append = true;
  bool partial = false;
  // graph mode (omx_attn_decode_fused_dynamic): the position is read from device memory by the kernel;
  // K/V views span max_rows rows, `position` above is 0, scratch is owned by the cache
  const int* pos_dev = nullptr;
  int max_rows = 0;
  void* scratch = nullptr;  // split-K partials + arrival counters, zero-initialised, fixed address
  size_t scratch_bytes = 0;
  // paged cache (paged_kv.cu): K/V views describe the page pool, per-sequence lengths live in device memory
  const struct PagedRef* paged = nullptr;
  // rows [0, stable_rows) of k / v are older than the stream's previous kernel (kv_cache_stable_rows): the TMA
  // kernel requests them before the dependency wait of its programmatic launch.  0 = unknown.
  int stable_rows = 0;
};
struct PagedRef {
  const int* block_table = nullptr;  // [B][bt_stride] page ids, device
  int bt_stride = 0;                 // pages per sequence the table can hold
  const int* lens_in = nullptr;      // [B] rows stored before the step (< 0: inactive slot), device
  int* lens_out = nullptr;           // [B] receives lens_in + 1 for the sequences that appended
  int64_t n_pages = 0;
};
// Bytes of cache-owned scratch a dynamic-position launch of this shape can need.
size_t decode_graph_scratch_bytes(int B, int Hkv, int Hq, int D, int dtype, int max_rows);
// q [B,Hq,1,D]; k/v views over Lk rows (Lk INCLUDES the new row when fused: the kernel reads
// rows [0, Lk-1) from memory and takes row Lk-1 from k_new/v_new, writing it to k/v as well).
bool decode_supported(const SdpaArgs& a, const char** why);
void decode_attention(const SdpaArgs& a, const DecodeFused& f, cudaStream_t stream);
// One thread per rank spins (bounded) until flags[r] has reached `expected` for every r.
void peer_wait(const unsigned* flags, int world, unsigned expected, int rank, cudaStream_t stream);
// Sequence-sharded decode, exchange step: waits (bounded) until flags[r] >= expected for every rank (flags may
// be null: no wait), then out[b,h,:] = sum_r w_r O_r / sum_r w_r with w_r = l_r 2^(m_r - max m) over the `world`
// float32 partial slots [world][B][Hq][D + 2].
void seqshard_merge(const omx_array* out, const float* partial, int world, int B, int Hq, int D,
                    const unsigned* flags, unsigned expected, int rank, cudaStream_t stream);

// ---- paged_kv.cu ----
struct PagedKVImpl;
PagedKVImpl* paged_create(int B, int H, int Dk, int Dv, int dtype, int64_t n_pages, int max_pages_per_seq);
void paged_destroy(PagedKVImpl* c);
int paged_offset(const PagedKVImpl* c);        // longest sequence
const int* paged_lengths(const PagedKVImpl* c);  // host mirror, [B]; -1 = released slot
int paged_batch(const PagedKVImpl* c);
int64_t paged_free_pages(const PagedKVImpl* c);
void paged_shape(const PagedKVImpl* c, int* B, int* H, int* Dk, int* Dv, int* dtype, int64_t* n_pages, int* max_pages);
void paged_reset(PagedKVImpl* c, int slot, bool deactivate, cudaStream_t stream);
void paged_reserve(PagedKVImpl* c, int rows_ahead, cudaStream_t stream);
void paged_append(PagedKVImpl* c, int slot0, const omx_array* keys, const omx_array* values, cudaStream_t stream);
void paged_materialize(PagedKVImpl* c, omx_array* keys_out, omx_array* values_out, cudaStream_t stream);
void paged_begin_step(PagedKVImpl* c, int n_q_heads, omx_array* kpool_view, omx_array* vpool_view, PagedRef* ref,
                      void** scratch, size_t* scratch_bytes, int* max_len_after, int* table_rows, cudaStream_t stream);
void paged_sync_lengths(PagedKVImpl* c, cudaStream_t stream);
void paged_end_step(PagedKVImpl* c);
void paged_trim(PagedKVImpl* c, int n, cudaStream_t stream);
void paged_pool_ptrs(const PagedKVImpl* c, void** kpool, void** vpool, const int** block_table);

// ---- sdpa_f32_tiled.cu ---- float32, Lq > 1, head_dim 64 / 128: 64 x 64 tiles on the FFMA pipe
bool sdpa_f32_tiled_supported(const SdpaArgs& a, const char** why);
void sdpa_f32_tiled(const SdpaArgs& a, cudaStream_t stream);

// ---- sdpa_mma.cu ---- bf16 / f16, keys up to 576 / values up to 512 features (multiples of 8): mma.sync tiles,
// the kv head's query group packed into the rows (absorbed MLA, head dims outside {64, 128})
bool sdpa_mma_supported(const SdpaArgs& a, const char** why);
void sdpa_mma(const SdpaArgs& a, cudaStream_t stream);
// Graph mode (single-token steps): the views in `a` span the pinned rows, the key count is *pos_dev + 1 (device);
// partials go to `scratch` (fixed address).  Same bits as sdpa_mma() on views of *pos_dev + 1 rows.
// `paged`: k / v in `a` are the page pools of the paged cache; per-sequence key counts and page ids come from device memory.
void sdpa_mma_dynamic(const SdpaArgs& a, const int* pos_dev, void* scratch, size_t scratch_bytes, cudaStream_t stream,
                      const struct PagedRef* paged = nullptr);
size_t sdpa_mma_graph_scratch_bytes(int B, int Hkv, int Hq, int Dv);
// single-token 16-bit calls outside head dim 128 with >= 2 query heads per kv head: faster here than on decode_simt
bool sdpa_mma_preferred_for_decode(const SdpaArgs& a);

// ---- fmha_sm100.cu ----
bool fmha_sm100_supported(const SdpaArgs& a, const char** why);
void fmha_sm100(const SdpaArgs& a, cudaStream_t stream);

}  // namespace omx
