/*
 * omx_attn.h -- C ABI of the B200-native attention hot path
 *               (RoPE -> KV-cache append/fetch -> scaled-dot-product attention).
 *
 * This is the drop-in boundary for ONE path of OminiX-MLX: the three calls every
 * model crate makes through mlx-rs-core / mlx-rs.  Each entry point names the
 * reference interface it replaces (paths relative to the OminiX-MLX checkout):
 *
 *   omx_fast_rope                          <- mlx_fast_rope
 *        mlx-rs/mlx-sys/src/mlx-c/mlx/c/fast.h:169-178 (bound at mlx-rs/src/fast.rs:31-45)
 *   omx_fast_rope_dynamic                  <- mlx_fast_rope_dynamic          fast.h:179-188
 *   omx_fast_scaled_dot_product_attention  <- mlx_fast_scaled_dot_product_attention
 *        fast.h:189-198 (bound at mlx-rs/src/fast.rs:138-150)
 *   omx_kv_cache_*                         <- KVCache (Rust logic over mlx_zeros /
 *        mlx_concatenate_axis / mlx_slice_update / mlx_slice): mlx-rs-core/src/cache.rs:92-195
 *   omx_concat_kv_cache_*                  <- ConcatKeyValueCache            cache.rs:45-85
 *   omx_attn_decode_fused                  <- the composite Attention::forward decode step
 *        qwen3-mlx/src/model.rs:186-212 (rope(q), rope(k), update_and_fetch, sdpa) in ONE launch
 *   omx_dit_rope / omx_dit_joint_attention / omx_dit_attn_fused <- FLUX.2-klein / Z-Image manual attention
 *        flux-klein-mlx/src/klein_model.rs:124-162,443-489,641-663; zimage-mlx/src/zimage_model.rs:208-235,345-388
 *   omx_set_error_handler / omx_last_error <- mlx_set_error_handler  mlx-c/mlx/c/error.h
 *
 * Differences from mlx-c that a binding must know:
 *   - EAGER, not lazy: work is enqueued on the given CUDA stream when the call is made
 *     (asynchronous w.r.t. the host, like async_eval); there is no graph and no eval().
 *   - Arrays are BORROWED DESCRIPTORS of device memory (pointer + shape + element strides),
 *     not owning handles.  The caller allocates outputs; only the KV cache owns memory.
 *   - sm_100a only, no CPU fallback: every entry point fails (status 1) without a B200.
 *
 * Conventions (same as mlx-c): every function returns 0 on success, 1 on error; no C++
 * exception crosses the boundary; the message goes to the registered handler, or, with no
 * handler registered, into a thread-local slot readable with omx_last_error().
 */
#ifndef OMX_ATTN_H
#define OMX_ATTN_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
/* the library is built with -fvisibility=hidden; everything declared here is exported */
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

#define OMX_ATTN_VERSION 100
#define OMX_MAX_NDIM 8
#define OMX_MAX_PEERS 8

/* Element types; numeric values are those of mlx_dtype (mlx-c/mlx/c/array.h:37-52). */
typedef enum omx_dtype_ {
  OMX_BOOL = 0,
  OMX_INT32 = 7,
  OMX_FLOAT16 = 9,
  OMX_FLOAT32 = 10,
  OMX_BFLOAT16 = 12
} omx_dtype;

/* Borrowed, strided view of device memory.  Strides are in ELEMENTS (a transposed
 * [B,L,H,D] -> [B,H,L,D] view or a [..., :offset, :] cache slice is passed as is). */
typedef struct omx_array_ {
  void* data;
  int32_t dtype; /* omx_dtype */
  int32_t ndim;
  int64_t shape[OMX_MAX_NDIM];
  int64_t strides[OMX_MAX_NDIM];
} omx_array;

/* mlx_optional_float (mlx-c/mlx/c/optional.h:32-35) */
typedef struct omx_optional_float_ {
  float value;
  bool has_value;
} omx_optional_float;

/* A cudaStream_t (NULL = the legacy default stream). Replaces mlx_stream. */
typedef void* omx_stream;

/* ---- errors ------------------------------------------------------------- */
typedef void (*omx_error_handler_func)(const char* msg, void* data);
void omx_set_error_handler(omx_error_handler_func handler, void* data, void (*dtor)(void*));
/* Last error message of the calling thread ("" if none); valid until the next failing call. */
const char* omx_last_error(void);
int omx_version(void);
/* 0 and *sm = 100 when the current device can run the kernels; 1 otherwise. */
int omx_device_check(int* sm);

/* ---- rope --------------------------------------------------------------- */
/*
 * out = rope(x): x [..., T, D] (ndim >= 3; T = axis -2), rotate the first `dims` features by
 * theta(t,i) = (offset + t) * scale * inv_freq[i]; traditional=false pairs (i, i+dims/2),
 * true pairs (2i, 2i+1); exactly one of base / freqs ([dims/2] float32) must be given.
 * Same position for every batch row.  out: same shape and dtype as x (any strides).
 * Numerics: cos/sin tables are computed on the HOST with libm (as the MLX CPU backend does),
 * cast to x's dtype, and the rotation rounds to x's dtype after every multiply/add.
 */
int omx_fast_rope(const omx_array* out, const omx_array* x, int dims, bool traditional,
                  omx_optional_float base, float scale, int offset,
                  const omx_array* freqs /* may be null */, omx_stream s);
/* offset: int32 scalar in DEVICE memory (CUDA-graph friendly). max_position bounds the table. */
int omx_fast_rope_dynamic(const omx_array* out, const omx_array* x, int dims, bool traditional,
                          omx_optional_float base, float scale,
                          const omx_array* offset, int max_position,
                          const omx_array* freqs /* may be null */, omx_stream s);

/* ---- rms_norm ----------------------------------------------------------- */
/*
 * out = rms_norm(x, weight, eps) over the last axis  <- mlx_fast_rms_norm (fast.h:163-168, bound at
 * mlx-rs/src/fast.rs:163-180; on this path: the per-head q_norm / k_norm, qwen3-mlx/src/model.rs:172-181).
 * weight: [D] in x's dtype, or null.  Same op order and rounding as the MLX CPU backend (f32 sum
 * left to right, 1/sqrt, round to dtype, then * weight), so downstream KV-cache bits are unchanged.
 */
int omx_fast_rms_norm(const omx_array* out, const omx_array* x, const omx_array* weight /* may be null */,
                      float eps, omx_stream s);

/* ---- scaled dot-product attention --------------------------------------- */
/*
 * out[B,Hq,Lq,Dv] = softmax(scale * q k^T + mask) v;  q [B,Hq,Lq,D], k [B,Hkv,Lk,D],
 * v [B,Hkv,Lk,Dv], Hq % Hkv == 0 (GQA: q head h reads kv head h / (Hq/Hkv), K/V not tiled).
 * mask_mode "" (mask_arr null: none; bool array: true = keep; float array: additive, same
 * dtype as q) or "causal" (lower-triangular, aligned bottom-right when Lq < Lk).
 * mask_arr broadcastable to [B,Hq,Lq,Lk].  sinks must be null (mlx-rs never passes any).
 * Softmax and accumulation in float32.
 */
int omx_fast_scaled_dot_product_attention(const omx_array* out, const omx_array* queries,
                                          const omx_array* keys, const omx_array* values,
                                          float scale, const char* mask_mode,
                                          const omx_array* mask_arr /* may be null */,
                                          const omx_array* sinks /* must be null */,
                                          omx_stream s);

/* ---- KV caches ---------------------------------------------------------- */
typedef struct omx_kv_cache_ {
  void* ctx;
} omx_kv_cache;

int omx_kv_cache_new(omx_kv_cache* res, int step /* KVCache::new(): 256 */);
int omx_kv_cache_free(omx_kv_cache c);
int omx_kv_cache_offset(omx_kv_cache c, int* offset);
int omx_kv_cache_reset(omx_kv_cache c); /* offset = 0; buffers untouched */
/*
 * keys [B,Hkv,n,Dk], values [B,Hkv,n,Dv] (any strides).  Appends at rows [offset, offset+n),
 * growing by ceil(n/step)*step zero rows (old buffer trimmed to `offset` first when
 * offset % step != 0) exactly as cache.rs:141-181; then returns VIEWS [.., :offset, :] of
 * the cache-owned buffers in keys_out / values_out.  A growth step moves the rows into a new buffer; the buffer it
 * replaces stays allocated (readable, with its old contents) until the NEXT growth or the cache's destruction, so a
 * fetched view survives one growth step -- e.g. a consumer still running on another stream -- but not two.
 */
int omx_kv_cache_update_and_fetch(omx_kv_cache c, const omx_array* keys, const omx_array* values,
                                  omx_array* keys_out, omx_array* values_out, omx_stream s);
/* Whole backing buffers [B,Hkv,cap,D] incl. the zero tail (cap = the reference's keys.shape[2]). */
int omx_kv_cache_state(omx_kv_cache c, omx_array* keys_buf, omx_array* values_buf);
/* Extension (mlx-lm's KVCache.trim; the reference only has the stub trim_cache,
 * mlx-rs-core/src/speculative.rs:165-167): drop the last min(n, offset) rows; buffers untouched. */
int omx_kv_cache_trim(omx_kv_cache c, int n, int* trimmed);
/* Extension: pre-size the device allocation (rows) so appends never reallocate.  Does not
 * change the logical capacity the reference rule produces. */
int omx_kv_cache_reserve(omx_kv_cache c, int rows);

int omx_concat_kv_cache_new(omx_kv_cache* res);
int omx_concat_kv_cache_free(omx_kv_cache c);
int omx_concat_kv_cache_offset(omx_kv_cache c, int* offset);
int omx_concat_kv_cache_update_and_fetch(omx_kv_cache c, const omx_array* keys,
                                         const omx_array* values, omx_array* keys_out,
                                         omx_array* values_out, omx_stream s);

/* ---- fused decode step -------------------------------------------------- */
/*
 * One launch for the decode step of Attention::forward (L == 1):
 *   off = cache.offset; q' = rope(q, off); k' = rope(k_new, off);
 *   cache.update_and_fetch(k', v_new); out = sdpa(q', K[:off+1], V[:off+1], sm_scale, none)
 * q [B,Hq,1,D], k_new/v_new [B,Hkv,1,D], out [B,Hq,1,D].  rope_dims == 0 skips the rotation.
 * keys_out / values_out (may be null) receive the fetched views.  Cache contents are
 * bit-identical to omx_fast_rope + omx_kv_cache_update_and_fetch.
 */
int omx_attn_decode_fused(const omx_array* out, const omx_array* q, const omx_array* k_new,
                          const omx_array* v_new, omx_kv_cache cache, int rope_dims,
                          bool traditional, omx_optional_float base, float rope_scale,
                          const omx_array* freqs /* may be null */, float sm_scale,
                          omx_array* keys_out, omx_array* values_out, omx_stream s);

/*
 * The same step with the per-head RMSNorm prologue of the Qwen3-family crates folded in
 * (qwen3-mlx/src/model.rs:172-212: q_norm(q), k_norm(k), rope, rope, update_and_fetch, sdpa):
 *   q' = rope(rms_norm(q, q_norm_weight, norm_eps), off); k' = rope(rms_norm(k_new, k_norm_weight, norm_eps), off)
 * q_norm_weight / k_norm_weight: [D] in q's dtype (either may be null = no norm on that side).
 */
int omx_attn_decode_fused_norm(const omx_array* out, const omx_array* q, const omx_array* k_new,
                               const omx_array* v_new, omx_kv_cache cache,
                               const omx_array* q_norm_weight, const omx_array* k_norm_weight,
                               float norm_eps, int rope_dims, bool traditional,
                               omx_optional_float base, float rope_scale,
                               const omx_array* freqs /* may be null */, float sm_scale,
                               omx_array* keys_out, omx_array* values_out, omx_stream s);

/*
 * The prefill spelling of the same composite (L >= 1 new tokens; Attention::forward with L > 1,
 * qwen3-mlx/src/model.rs:172-212): k' = rope(norm(k_new)) is written straight into the cache rows
 * [offset, offset+L) (no intermediate k' array, no second copy), v_new is copied into its rows, q' goes
 * to library scratch, then attention over the fetched K/V with the caller's mask rule
 * (mask_mode / mask_arr exactly as omx_fast_scaled_dot_product_attention: "causal", a bool array from
 * create_causal_mask, an additive array, or none).  Cache bits = the unfused op chain's.
 */
int omx_attn_prefill_fused(const omx_array* out, const omx_array* q, const omx_array* k_new,
                           const omx_array* v_new, omx_kv_cache cache,
                           const omx_array* q_norm_weight /* may be null */,
                           const omx_array* k_norm_weight /* may be null */, float norm_eps,
                           int rope_dims, bool traditional, omx_optional_float base, float rope_scale,
                           const omx_array* freqs /* may be null */, float sm_scale,
                           const char* mask_mode, const omx_array* mask_arr /* may be null */,
                           omx_array* keys_out, omx_array* values_out, omx_stream s);

/* ---- CUDA-graph decode loop -------------------------------- */
/*
 * The reference hides per-op launch cost behind MLX's lazy graph + async_eval double buffering
 * (qwen3-mlx/src/model.rs:798-844: Generate::next evaluates token t while building t+1).  The
 * eager equivalent on CUDA is to capture the per-layer launches of one decode step ONCE and replay
 * the graph per token.  A captured launch cannot take the position as a host argument, so:
 *   - omx_kv_cache_prepare_graph pins the cache buffers (and a cache-owned split-K scratch) for
 *     positions [0, max_rows) -- addresses then stay fixed until the cache grows past max_rows or
 *     prepare_graph is called again (re-capture after either).  n_q_heads = Hq of the launches;
 *   - omx_attn_decode_fused_dynamic is omx_attn_decode_fused_norm with `off` read by the KERNEL
 *     from *position (device int32, shared by all layers of the model): row *position is written,
 *     keys [0, *position] are attended.  It does NOT touch the host-side offset;
 *   - omx_device_counter_add bumps the device position (one 1-thread launch per token step);
 *   - omx_kv_cache_advance(c, n) is the host bookkeeping after n such steps (offset += n, logical
 *     capacity by the cache.rs:141-181 rule, no copy), so offset()/state()/update_and_fetch agree
 *     with what the launches did.
 * Every launch is capture-safe once the same call has run eagerly (rope table and scratch exist).
 * Same kernels and arithmetic as omx_attn_decode_fused_norm: the appended cache rows are always
 * bit-identical, the outputs are bit-identical whenever the split-K plan coincides (it is fixed at
 * capture from max_rows instead of following the current length) and differ by f32 summation
 * order otherwise.
 */
int omx_kv_cache_prepare_graph(omx_kv_cache c, int max_rows, int n_q_heads, omx_stream s);
int omx_kv_cache_advance(omx_kv_cache c, int n, omx_stream s);
int omx_attn_decode_fused_dynamic(const omx_array* out, const omx_array* q, const omx_array* k_new,
                                  const omx_array* v_new, omx_kv_cache cache,
                                  const omx_array* q_norm_weight /* may be null */,
                                  const omx_array* k_norm_weight /* may be null */, float norm_eps,
                                  int rope_dims, bool traditional, omx_optional_float base,
                                  float rope_scale, float sm_scale,
                                  const int32_t* position /* device */, omx_stream s);
int omx_device_counter_add(int32_t* counter /* device */, int delta, omx_stream s);

/* ---- paged KV cache (north_star: "fuses RoPE and the paged KV append") --------- */
/*
 * The reference's KVCache (mlx-rs-core/src/cache.rs:92-195) concatenates a fresh zero block every 256
 * tokens -- an O(S) copy per growth -- and keeps one rectangular [B,Hkv,cap,D] buffer, so every sequence of
 * a batch has the same length.  The paged cache keeps the trait contract (offset / update_and_fetch / reset,
 * rows appended bit for bit) over a page pool:
 *   pool K, pool V  [n_pages][Hkv][64][D], allocated ONCE; growth = a page id from the free list (zero copy)
 *   block table     int32 [batch][max_pages_per_seq]: page of rows [64 t, 64 t + 64) of each sequence
 *   lengths         per sequence: ragged batches, slots that are released and reused (continuous batching)
 * One page = one 64-key pipeline stage of the decode kernel = one TMA box pair per tensor.
 *   omx_paged_kv_cache_update_and_fetch  KeyValueCache::update_and_fetch for ALL sequences ([B,Hkv,n,D] appended
 *        to each sequence at its own length).  keys_out / values_out may be null; when given, the reference's
 *        [B,Hkv,offset,D] views (cache.rs:190-193) are MATERIALISED into a cache-owned contiguous buffer
 *        (offset = the longest sequence; rows past a shorter sequence's end read +0.0) -- valid until the next
 *        materialising call.
 *   omx_paged_kv_cache_append_slot       the same for one sequence ([1,Hkv,n,D]): ragged prefill
 *   omx_paged_kv_cache_fetch             materialise without appending
 *   omx_paged_kv_cache_reset             KeyValueCache::reset for one sequence (slot) or all (slot < 0):
 *        length 0, pages back to the free list
 *   omx_paged_kv_cache_release           mark a slot inactive: omx_attn_decode_fused_paged skips it
 *   omx_paged_kv_cache_reserve           pre-assign pages for `rows_ahead` more rows per active sequence: the next
 *        rows_ahead fused steps need no host-side allocation (capturable into a CUDA graph, an even number
 *        of steps per capture -- the device lengths are double-buffered by step parity)
 *   omx_paged_kv_cache_sync_lengths      host mirror <- device lengths (synchronises the stream): after capturing /
 *        replaying fused steps, which advance the device lengths without the host seeing it
 *   omx_paged_kv_cache_trim              drop the last n rows of every active sequence (lengths only)
 *   omx_attn_decode_fused_paged          omx_attn_decode_fused_norm over the pages, ONE launch: per sequence b,
 *        off = len[b] (read by the kernel from device memory); q' = rope(norm(q[b]), off); k' = rope(norm(k_new[b]),
 *        off) stored with v_new[b] into row off % 64 of page block_table[b][off / 64]; attention of q' over the
 *        sequence's off + 1 keys, K/V tiles fetched page by page with TMA; len[b] += 1.  Appended rows and
 *        outputs equal the contiguous cache's (same kernels, same arithmetic).
 */
typedef struct omx_paged_kv_cache_ {
  void* ctx;
} omx_paged_kv_cache;
int omx_paged_kv_cache_new(omx_paged_kv_cache* res, int batch, int n_kv_heads, int head_dim_k, int head_dim_v,
                           int dtype /* mlx_dtype value */, int64_t n_pages, int max_pages_per_seq);
int omx_paged_kv_cache_free(omx_paged_kv_cache c);
int omx_paged_kv_cache_offset(omx_paged_kv_cache c, int* offset /* longest sequence */);
int omx_paged_kv_cache_lengths(omx_paged_kv_cache c, int32_t* lens /* host, [batch]; -1 = released slot */);
int omx_paged_kv_cache_free_pages(omx_paged_kv_cache c, int64_t* n);
int omx_paged_kv_cache_reset(omx_paged_kv_cache c, int slot, omx_stream s);
int omx_paged_kv_cache_release(omx_paged_kv_cache c, int slot, omx_stream s);
int omx_paged_kv_cache_reserve(omx_paged_kv_cache c, int rows_ahead, omx_stream s);
int omx_paged_kv_cache_sync_lengths(omx_paged_kv_cache c, omx_stream s);
int omx_paged_kv_cache_trim(omx_paged_kv_cache c, int n, omx_stream s);
int omx_paged_kv_cache_update_and_fetch(omx_paged_kv_cache c, const omx_array* keys, const omx_array* values,
                                        omx_array* keys_out /* may be null */, omx_array* values_out /* may be null */,
                                        omx_stream s);
int omx_paged_kv_cache_append_slot(omx_paged_kv_cache c, int slot, const omx_array* keys, const omx_array* values,
                                   omx_stream s);
int omx_paged_kv_cache_fetch(omx_paged_kv_cache c, omx_array* keys_out, omx_array* values_out, omx_stream s);
/* Introspection (tests, debuggers): pool bases, the HOST mirror of the block table ([batch][max_pages_per_seq],
 * valid until the next call on the cache) and its row length. */
int omx_paged_kv_cache_pages(omx_paged_kv_cache c, void** k_pool, void** v_pool, const int32_t** block_table,
                             int* max_pages_per_seq);
int omx_attn_decode_fused_paged(const omx_array* out, const omx_array* q, const omx_array* k_new,
                                const omx_array* v_new, omx_paged_kv_cache cache,
                                const omx_array* q_norm_weight /* may be null */,
                                const omx_array* k_norm_weight /* may be null */, float norm_eps, int rope_dims,
                                bool traditional, omx_optional_float base, float rope_scale, float sm_scale,
                                omx_stream s);

/* ---- head-sharded single-sequence decode (BASELINE C5) -------------------- */
/*
 * The reference has no multi-device path (MLX is single-GPU); this is the exchange step the
 * kv-head-sharded layout of SURVEY 8(e) needs.  Rank r owns kv heads [r*Hkv/world, ...) and the
 * q heads that read them; every rank must end the step holding the full [B,Hq,1,D] output.
 * Instead of a separate all-gather, the decode kernel's final store writes this rank's head slice
 * straight into EVERY rank's output buffer through NVLink peer mappings, and the last CTA of the
 * launch bumps one arrival counter per rank (system-scope fence, then atomic).  omx_peer_wait
 * enqueues a one-warp kernel that returns once all `world` counters of the local rank reached
 * `expected` (= number of sharded steps issued so far).  expected == 0 means "as many arrivals as this
 * rank has itself signalled" (its own launch, earlier on the stream, bumped its own counter): no host-side
 * step count, so launch + wait can be captured once into a CUDA graph and replayed; same for
 * omx_seqshard_merge.
 *
 * BUFFER CONTRACT: the wait for step t proves that every peer has STORED its step-t slice, not that the
 * peers have finished READING their own step-t buffer.  The caller must therefore alternate between TWO
 * output buffers by step parity (out[] of the peer group shifted accordingly, as for the sequence-sharded
 * partials below) and consume step t's buffer on the same stream before it launches step t+1: a rank
 * cannot start step t+2 (which reuses step t's buffer) before its own wait for t+1 has returned, i.e.
 * before every peer has launched t+1, stream-ordered after that peer's reads of step t.  With a single
 * buffer a fast rank's step t+1 would overwrite a slow rank's step-t output while it is being read.
 * (ominix-mlx_b200/parallel.py HeadShardedDecode does this; tests/test_parallel_gpu.py delays one rank.)
 */
typedef struct omx_peer_group_ {
  int32_t world; /* <= OMX_MAX_PEERS */
  int32_t rank;
  void* out[OMX_MAX_PEERS];       /* rank r's FULL output buffer, mapped into this process (out[rank] = local) */
  uint32_t* flags[OMX_MAX_PEERS]; /* rank r's uint32[world] arrival counters (zero-initialised), mapped likewise */
} omx_peer_group;
/* Same contract as omx_attn_decode_fused for the LOCAL heads (q [B,Hq_local,1,D], k_new/v_new
 * [B,Hkv_local,1,D], cache = this rank's shard).  out_full describes the full [B,Hq_total,1,D]
 * buffer (same strides on every rank); rows [head_offset, head_offset + Hq_local) are written on
 * every rank. */
int omx_attn_decode_fused_sharded(const omx_array* out_full, const omx_array* q,
                                  const omx_array* k_new, const omx_array* v_new,
                                  omx_kv_cache cache, int rope_dims, bool traditional,
                                  omx_optional_float base, float rope_scale,
                                  const omx_array* freqs /* may be null */, float sm_scale,
                                  const omx_peer_group* peers, int head_offset, omx_stream s);
int omx_peer_wait(const omx_peer_group* peers, uint32_t expected, omx_stream s);
/* The same step with the wait folded into the launch: the CTA that publishes this rank's arrival then spins
 * (bounded) until every peer's arrival of the same step has reached the local counters, so the launch completes
 * only when the full [B,Hq_total,1,D] output is in place -- ONE launch per step, no omx_peer_wait.  Every rank of
 * the group must launch its step (a kernel only ever waits on its own SMs; no rank depends on another rank's
 * launch having been scheduled first). */
int omx_attn_decode_fused_sharded_sync(const omx_array* out_full, const omx_array* q,
                                       const omx_array* k_new, const omx_array* v_new,
                                       omx_kv_cache cache, int rope_dims, bool traditional,
                                       omx_optional_float base, float rope_scale,
                                       const omx_array* freqs /* may be null */, float sm_scale,
                                       const omx_peer_group* peers, int head_offset, omx_stream s);

/*
 * The same step with a DATA + FLAG exchange (what NCCL calls the LL protocol): each rank owns a staging buffer of
 * 8-byte words {payload (two 16-bit or one 32-bit element), sequence number of the step}; the thread that holds
 * final output values stores them as such words into every peer's staging buffer (an aligned 8-byte store is one
 * NVLink transaction, so a word whose flag equals the step's sequence number carries the step's payload) and
 * then polls its own staging buffer for the same words of every peer, writing them into the local out_full.
 * No system-scope fence, no arrival counters, no last-CTA ticket: the exchange costs one NVLink store latency.
 * The launch completes when the full [B,Hq_total,1,D] output is in place in the LOCAL, private out_full (peers
 * never write it, so it needs no double buffering).
 *   staging[r]: rank r's buffer, uint64 [2][world][B * Hq_local * D * sizeof(T) / 4] (two halves alternate by
 *               step parity: a rank can run at most one step ahead of a peer), zero-initialised, mapped into
 *               this process; seq: LOCAL uint32 counter of completed steps (zero-initialised; every rank of the
 *               group runs the same sequence of sharded launches, so the counters agree).
 * Every rank of the group must launch its step.  Launch shapes that do not finish in the all-CTA combine run the
 * plain local step followed by a one-CTA exchange kernel with the same protocol (two launches).
 */
typedef struct omx_ll_group_ {
  int32_t world; /* <= OMX_MAX_PEERS */
  int32_t rank;
  void* staging[OMX_MAX_PEERS];
  uint32_t* seq;
} omx_ll_group;
int omx_attn_decode_fused_sharded_ll(const omx_array* out_full, const omx_array* q, const omx_array* k_new,
                                     const omx_array* v_new, omx_kv_cache cache, int rope_dims, bool traditional,
                                     omx_optional_float base, float rope_scale,
                                     const omx_array* freqs /* may be null */, float sm_scale,
                                     const omx_ll_group* group, int head_offset, omx_stream s);
/* bytes of one rank's staging buffer for a [B,Hq_local,1,D] slice of `dtype` */
size_t omx_ll_staging_bytes(int world, int64_t B, int64_t Hq_local, int64_t D, int dtype);

/* ---- sequence-sharded single-sequence decode ---------------- */
/*
 * For contexts that should not (or do not) live on one GPU: rank r keeps the K/V rows of the token
 * positions it owns (any assignment: attention is a sum over keys; the host mirror uses position % world)
 * for ALL heads.  One step = one decode launch per rank + one exchange:
 *   omx_attn_decode_seqshard: q' = rope(q, position); on the rank with append = true also
 *     k' = rope(k_new, position) appended to the LOCAL cache row offset; attention of q' over the local rows;
 *     the result -- the normalised output of the local keys and their (m, l) in the log2 domain, float32
 *     [B,Hq,D+2] -- is stored into slot `rank` of EVERY rank's partial buffer [world,B,Hq,D+2] through the
 *     peer mappings (peers->out[r] = rank r's buffer; null peers = single rank), arrival counters as in
 *     omx_attn_decode_fused_sharded.  `position` is the GLOBAL position of the new token.
 *   omx_seqshard_merge: waits (bounded) for all `world` arrivals of step `expected`, then
 *     out[b,h,:] = sum_r w_r O_r / sum_r w_r,  w_r = l_r 2^(m_r - max_r m_r)      (one launch).
 * The partial buffer must be double-buffered by step parity by the caller (a rank may run one step ahead).
 */
int omx_attn_decode_seqshard(const omx_array* partial /* float32 [world,B,Hq,D+2], this rank's buffer */,
                             const omx_array* q, const omx_array* k_new /* null unless append */,
                             const omx_array* v_new /* null unless append */, omx_kv_cache cache,
                             int rope_dims, bool traditional, omx_optional_float base, float rope_scale,
                             int position, bool append, float sm_scale,
                             const omx_peer_group* peers /* may be null */, omx_stream s);
int omx_seqshard_merge(const omx_array* out /* [B,Hq,1,D] */, const omx_array* partial,
                       const omx_peer_group* peers /* may be null: no wait */, uint32_t expected,
                       omx_stream s);

/* ---- DiT joint attention ------------------------------------------------ */
/* Table-driven interleaved rope: x [B,S,H,D], cos/sin [B,S,D/2] (x's dtype);
 * out0 = x0*c - x1*s, out1 = x1*c + x0*s per adjacent pair, rounding after every op. */
int omx_dit_rope(const omx_array* out, const omx_array* x, const omx_array* cos,
                 const omx_array* sin, omx_stream s);
/* Non-causal joint attention over the concatenated [txt; img] sequence.  q/k/v/out are
 * [B,H,S,D] views (pass the [B,S,H,D] storage with transposed strides); add_mask optional
 * float32 [Lq,Lk].  out dtype may be q's dtype or float32 (the reference chain promotes). */
int omx_dit_joint_attention(const omx_array* out, const omx_array* q, const omx_array* k,
                            const omx_array* v, float scale, const omx_array* add_mask,
                            omx_stream s);

/*
 * One DiT attention block's attention, prologue included (klein_model.rs:443-489 double stream,
 * :641-663 single stream; zimage_model.rs:345-388): for each of the n_streams (1 or 2; FLUX's
 * [txt, img] in that order) q_i, k_i, v_i [B,S_i,H,D] are normalised per head (RMSNorm weights
 * q_norm_weight[i] / k_norm_weight[i], [D]; the list or an entry may be null), rotated with rows
 * [sum S_<i, ...) of the per-token tables cos / sin [B,S_total,D/2] (null = no rotation) and
 * concatenated along the sequence -- ONE launch for all of it -- then
 * out[B,S_total,H,Dv] = softmax(scale q k^T [+ add_mask]) v over the joint sequence (k/v may carry
 * fewer heads than q: Z-Image's repeat_axis GQA).  The caller slices out[:, :S_0] / out[:, S_0:].
 */
int omx_dit_attn_fused(const omx_array* out, int n_streams, const omx_array* const* q,
                       const omx_array* const* k, const omx_array* const* v,
                       const omx_array* const* q_norm_weight /* may be null */,
                       const omx_array* const* k_norm_weight /* may be null */, float norm_eps,
                       const omx_array* cos /* may be null */, const omx_array* sin /* may be null */,
                       float scale, const omx_array* add_mask /* may be null */, omx_stream s);

/* ---- introspection (tests / bench) -------------------------------------- */
/* Name of the kernel family the last successful attention call on this thread dispatched to
 * ("decode_hmma_tma", "decode_simt", "fmha_tcgen05", "sdpa_mma", "sdpa_f32_tiled", "sdpa_generic", ...). */
const char* omx_last_kernel(void);
/* Number of kernel launches issued by this library on the calling thread since the last reset. */
int64_t omx_launch_count(bool reset);
/* Force a kernel family for subsequent attention calls on this thread (NULL/"" = auto). */
int omx_force_kernel(const char* name);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* OMX_ATTN_H */
