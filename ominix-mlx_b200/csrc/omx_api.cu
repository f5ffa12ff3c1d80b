// omx_api.cu -- the extern "C" boundary of libomx_attn: validation, dispatch, error routing.
//
// Mirrors the conventions of mlx-c (mlx-c/mlx/c/fast.cpp:544-633, error.cpp:12-53): every entry
// point returns 0/1, exceptions never cross the ABI, messages go to the registered handler or a
// thread-local slot.  Validation messages follow MLX's so that the reference's callers see the
// same failures (rank, head-dim, GQA divisibility, dtype, mask broadcast / promotion).
#include <cstdarg>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "omx_common.cuh"
#include "omx_internal.h"

namespace omx {

namespace {

thread_local std::string t_last_error;
thread_local std::string t_last_kernel;
thread_local std::string t_forced_kernel;
thread_local int64_t t_launches = 0;

std::mutex g_handler_mu;
omx_error_handler_func g_handler = nullptr;
void* g_handler_data = nullptr;
void (*g_handler_dtor)(void*) = nullptr;

void report(const char* msg) {
  t_last_error = msg;
  std::lock_guard<std::mutex> lk(g_handler_mu);
  if (g_handler) g_handler(msg, g_handler_data);
}

template <typename F>
int guarded(F&& f) {
  try {
    f();
    return 0;
  } catch (const std::exception& e) {
    report(e.what());
    return 1;
  } catch (...) {
    report("unknown error");
    return 1;
  }
}

struct Scratch {
  void* p = nullptr;
  size_t bytes = 0;
};
std::mutex g_ws_mu;
std::map<std::pair<int, cudaStream_t>, Scratch> g_ws, g_ws_outer, g_ctr;

void* grow(std::map<std::pair<int, cudaStream_t>, Scratch>& pool, size_t bytes, cudaStream_t stream,
           size_t min_bytes) {
  int dev = 0;
  OMX_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(g_ws_mu);
  Scratch& s = pool[{dev, stream}];
  if (s.bytes < bytes) {
    size_t nb = std::max(bytes, std::max(min_bytes, 2 * s.bytes));
    void* np = nullptr;
    OMX_CUDA(cudaMallocAsync(&np, nb, stream));
    OMX_CUDA(cudaMemsetAsync(np, 0, nb, stream));
    if (s.p) OMX_CUDA(cudaFreeAsync(s.p, stream));  // stream-ordered: earlier kernels finish first
    s.p = np;
    s.bytes = nb;
  }
  return s.p;
}

int g_sm_count[64] = {0};

}  // namespace

void note_launch(const char* family) { t_last_kernel = family; }
void count_launch() { ++t_launches; }

void* get_workspace(size_t bytes, cudaStream_t stream) { return grow(g_ws, bytes, stream, 1 << 20); }
void* get_outer_workspace(size_t bytes, cudaStream_t stream) { return grow(g_ws_outer, bytes, stream, 1 << 20); }
int* get_counters(size_t count, cudaStream_t stream) {
  return (int*)grow(g_ctr, count * sizeof(int), stream, 1 << 16);
}

void* get_tagged_workspace(size_t bytes, cudaStream_t stream, unsigned** seq) {
  static std::map<std::pair<int, cudaStream_t>, Scratch> pool, seqs;
  static std::map<std::pair<int, cudaStream_t>, uint64_t> launches;
  unsigned* sq = (unsigned*)grow(seqs, 256, stream, 256);
  void* ws = grow(pool, bytes, stream, 1 << 20);  // (a new, zeroed pool under an old counter: stale tags are all 0)
  int dev = 0;
  OMX_CUDA(cudaGetDevice(&dev));
  {
    std::lock_guard<std::mutex> lk(g_ws_mu);
    // the 32-bit tag wraps after 4 G launches: long before that, start over from a clean pool (stream-ordered)
    if ((++launches[{dev, stream}] & ((1ull << 30) - 1)) == 0) {
      const Scratch& sc = pool[{dev, stream}];
      OMX_CUDA(cudaMemsetAsync(sc.p, 0, sc.bytes, stream));
      OMX_CUDA(cudaMemsetAsync(sq, 0, 256, stream));
    }
  }
  *seq = sq;
  return ws;
}

int sm_count() {
  int dev = 0;
  OMX_CUDA(cudaGetDevice(&dev));
  if (dev < 64 && g_sm_count[dev]) return g_sm_count[dev];
  int n = 0;
  OMX_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  if (dev < 64) g_sm_count[dev] = n;
  return n;
}

namespace {

void require_device() {
  int dev = 0;
  OMX_CUDA(cudaGetDevice(&dev));
  static thread_local int checked_dev = -1;
  if (checked_dev == dev) return;
  int major = 0;
  OMX_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  OMX_CHECK(major == 10, "libomx_attn is built for sm_100a (B200) only; device %d has compute capability %d.x "
                         "and there is no fallback path", dev, major);
  checked_dev = dev;
}

void check_arr(const omx_array* a, const char* name, int ndim) {
  OMX_CHECK(a != nullptr, "[omx] %s is null", name);
  OMX_CHECK(a->ndim == ndim, "[scaled_dot_product_attention] %s must be %d-dimensional, got %d", name, ndim,
            a->ndim);
}

// Fill SdpaArgs from the raw arrays with MLX's validation rules.
SdpaArgs make_sdpa_args(const omx_array* out, const omx_array* q, const omx_array* k, const omx_array* v,
                        float scale, const char* mask_mode, const omx_array* mask_arr,
                        const omx_array* sinks) {
  OMX_CHECK(q && k && v && out, "[scaled_dot_product_attention] null array");
  for (const omx_array* t : {q, k, v}) {
    OMX_CHECK(t->ndim == 4,
              "[scaled_dot_product_attention] input with shape of %d dims is not supported; expected "
              "[B, N, T, D]", t->ndim);
  }
  OMX_CHECK(sinks == nullptr || sinks->data == nullptr,
            "[scaled_dot_product_attention] attention sinks are not supported");
  OMX_CHECK(q->shape[0] == k->shape[0] && q->shape[0] == v->shape[0],
            "[scaled_dot_product_attention] mismatching batch dimension for input");
  OMX_CHECK(q->shape[3] == k->shape[3],
            "[scaled_dot_product_attention] query, keys expected to have matching last dimension; found "
            "%lld and %lld", (long long)q->shape[3], (long long)k->shape[3]);
  OMX_CHECK(k->shape[1] == v->shape[1],
            "[scaled_dot_product_attention] keys, values expected to have matching n_kv_heads; found %lld "
            "and %lld", (long long)k->shape[1], (long long)v->shape[1]);
  OMX_CHECK(k->shape[2] == v->shape[2],
            "[scaled_dot_product_attention] keys, values expected to have matching sequence length");
  OMX_CHECK(k->shape[1] > 0 && q->shape[1] % k->shape[1] == 0,
            "[scaled_dot_product_attention] n_heads must be a multiple of n_kv_heads, found n_heads %lld "
            "for n_kv_heads %lld", (long long)q->shape[1], (long long)k->shape[1]);
  OMX_CHECK(is_float_dtype(q->dtype),
            "[scaled_dot_product_attention] Received unsupported type %s", dtype_name(q->dtype));
  OMX_CHECK(q->dtype == k->dtype && q->dtype == v->dtype,
            "[scaled_dot_product_attention] q, k, v must share one dtype (the reference promotes; this "
            "boundary does not)");
  SdpaArgs a{};
  a.out = out; a.q = q; a.k = k; a.v = v;
  a.scale = scale;
  a.B = (int)q->shape[0]; a.Hq = (int)q->shape[1]; a.Lq = (int)q->shape[2]; a.D = (int)q->shape[3];
  a.Hkv = (int)k->shape[1]; a.Lk = (int)k->shape[2]; a.Dv = (int)v->shape[3];
  OMX_CHECK(out->ndim == 4 && out->shape[0] == a.B && out->shape[1] == a.Hq && out->shape[2] == a.Lq &&
                out->shape[3] == a.Dv,
            "[scaled_dot_product_attention] out must be [B, n_heads, L_q, D_v]");
  OMX_CHECK(out->dtype == q->dtype || out->dtype == OMX_FLOAT32,
            "[scaled_dot_product_attention] out dtype must be the input dtype (or float32)");
  const std::string mode = mask_mode ? mask_mode : "";
  const bool has_arr = mask_arr && mask_arr->data;
  if (mode == "causal") {
    OMX_CHECK(!has_arr, "[scaled_dot_product_attention] Invalid mask_arrs for mask_mode 'causal'. No array "
                        "masks supported.");
    a.mask_mode = MASK_CAUSAL;
  } else if (mode.empty() || mode == "array") {
    if (has_arr) {
      OMX_CHECK(mask_arr->ndim <= 4, "[scaled_dot_product_attention] the mask with shape of %d dims is not "
                                     "supported", mask_arr->ndim);
      if (mask_arr->dtype == OMX_BOOL) {
        a.mask_mode = MASK_BOOL;
      } else {
        OMX_CHECK(is_float_dtype(mask_arr->dtype) &&
                      (mask_arr->dtype == q->dtype || (mask_arr->dtype == OMX_FLOAT32 && out->dtype == OMX_FLOAT32)),
                  "[scaled_dot_product_attention] Mask type must promote to output type %s.",
                  dtype_name(q->dtype));
        a.mask_mode = MASK_ADD;
      }
      a.mask = mask_arr;
      const int64_t full[4] = {a.B, a.Hq, a.Lq, a.Lk};
      const int lead = 4 - mask_arr->ndim;
      for (int i = 0; i < 4; ++i) {
        if (i < lead) {
          a.mask_strides[i] = 0;
          continue;
        }
        const int64_t n = mask_arr->shape[i - lead];
        OMX_CHECK(n == full[i] || n == 1,
                  "[scaled_dot_product_attention] Mask with shape axis %d = %lld is not broadcastable to "
                  "[%d, %d, %d, %d]", i - lead, (long long)n, a.B, a.Hq, a.Lq, a.Lk);
        a.mask_strides[i] = (n == 1 && full[i] != 1) ? 0 : mask_arr->strides[i - lead];
      }
    } else {
      a.mask_mode = MASK_NONE;
    }
  } else {
    OMX_CHECK(false, "[scaled_dot_product_attention] Invalid mask_mode %s. mask_mode must be 'causal', "
                     "'array' or ''.", mode.c_str());
  }
  return a;
}

void dispatch_sdpa(const SdpaArgs& a, cudaStream_t stream) {
  if ((int64_t)a.B * a.Hq * a.Lq * a.Dv == 0) return;
  const std::string& force = t_forced_kernel;
  const char* why = nullptr;
  if (force.empty() && sdpa_mma_preferred_for_decode(a)) {
    sdpa_mma(a, stream);  // grouped-query single-token calls outside head dim 128: key-group mma.sync tiles
    return;
  }
  if (force.empty() || force == "decode" || force == "decode_simt" || force == "decode_hmma_tma") {
    if (a.out->dtype == a.q->dtype && decode_supported(a, &why)) {
      DecodeFused none;
      decode_attention(a, none, stream);
      return;
    }
    OMX_CHECK(force.empty(), "forced kernel '%s' does not support this call: %s", force.c_str(), why ? why : "?");
  }
  if (force.empty() || force == "fmha_tcgen05") {
    // Small query blocks (speculative / chunked decode: 1 < Lq < 32 rows against a long cache) fill only part of
    // the 256-row CTA tile, but the tile streams K/V with TMA and is still memory-bound: measured at B4, 32 / 8
    // heads, 8192 keys, Lq = 2 .. 31: 97 us here vs 11 - 28 ms on the one-warp-per-row kernel
    // (scripts/gpu_r02_small_lq.py), so every multi-row call the kernel supports takes it.
    if (fmha_sm100_supported(a, &why) && (a.Lq >= 2 || !force.empty())) {
      fmha_sm100(a, stream);
      return;
    }
    OMX_CHECK(force.empty(), "forced kernel '%s' does not support this call: %s", force.c_str(), why ? why : "?");
  }
  if (force.empty() || force == "sdpa_mma") {
    // 16-bit calls the two kernels above refuse -- keys wider than values (absorbed MLA), head dims outside
    // {64, 128}, layouts they do not take: mma.sync tiles with the kv head's whole query group packed into the rows
    if (sdpa_mma_supported(a, &why)) {
      sdpa_mma(a, stream);
      return;
    }
    OMX_CHECK(force.empty(), "forced kernel '%s' does not support this call: %s", force.c_str(), why ? why : "?");
  }
  if (force.empty() || force == "sdpa_f32_tiled") {
    // float32 with more than a handful of query rows: shared-memory tiles instead of one warp per row
    if (sdpa_f32_tiled_supported(a, &why) && (a.Lq >= 16 || !force.empty())) {
      sdpa_f32_tiled(a, stream);
      return;
    }
    OMX_CHECK(force.empty(), "forced kernel '%s' does not support this call: %s", force.c_str(), why ? why : "?");
  }
  sdpa_generic(a, stream);
}

}  // namespace
}  // namespace omx

using namespace omx;

extern "C" {

void omx_set_error_handler(omx_error_handler_func handler, void* data, void (*dtor)(void*)) {
  std::lock_guard<std::mutex> lk(g_handler_mu);
  if (g_handler_dtor && g_handler_data) g_handler_dtor(g_handler_data);
  g_handler = handler;
  g_handler_data = data;
  g_handler_dtor = dtor;
}

const char* omx_last_error(void) { return t_last_error.c_str(); }
int omx_version(void) { return OMX_ATTN_VERSION; }

int omx_device_check(int* sm) {
  return guarded([&] {
    require_device();
    if (sm) *sm = 100;
  });
}

const char* omx_last_kernel(void) { return t_last_kernel.c_str(); }

int64_t omx_launch_count(bool reset) {
  const int64_t n = t_launches;
  if (reset) t_launches = 0;
  return n;
}

int omx_force_kernel(const char* name) {
  return guarded([&] {
    const std::string n = name ? name : "";
    OMX_CHECK(n.empty() || n == "decode" || n == "decode_simt" || n == "decode_hmma_tma" ||
                  n == "fmha_tcgen05" || n == "sdpa_generic" || n == "sdpa_f32_tiled" || n == "sdpa_mma",
              "unknown kernel family '%s'", n.c_str());
    t_forced_kernel = n;
  });
}

int omx_fast_rms_norm(const omx_array* out, const omx_array* x, const omx_array* weight, float eps, omx_stream s) {
  return guarded([&] {
    require_device();
    rms_norm_forward(out, x, weight, eps, (cudaStream_t)s);
  });
}

int omx_fast_rope(const omx_array* out, const omx_array* x, int dims, bool traditional,
                  omx_optional_float base, float scale, int offset, const omx_array* freqs, omx_stream s) {
  return guarded([&] {
    require_device();
    rope_forward(out, x, dims, traditional, base, scale, offset, nullptr, 0, freqs, (cudaStream_t)s);
  });
}

int omx_fast_rope_dynamic(const omx_array* out, const omx_array* x, int dims, bool traditional,
                          omx_optional_float base, float scale, const omx_array* offset, int max_position,
                          const omx_array* freqs, omx_stream s) {
  return guarded([&] {
    require_device();
    OMX_CHECK(offset != nullptr, "[rope] offset array is null");
    rope_forward(out, x, dims, traditional, base, scale, 0, offset, max_position, freqs, (cudaStream_t)s);
  });
}

int omx_fast_scaled_dot_product_attention(const omx_array* out, const omx_array* queries,
                                          const omx_array* keys, const omx_array* values, float scale,
                                          const char* mask_mode, const omx_array* mask_arr,
                                          const omx_array* sinks, omx_stream s) {
  return guarded([&] {
    require_device();
    SdpaArgs a = make_sdpa_args(out, queries, keys, values, scale, mask_mode, mask_arr, sinks);
    dispatch_sdpa(a, (cudaStream_t)s);
  });
}

int omx_kv_cache_new(omx_kv_cache* res, int step) {
  return guarded([&] {
    OMX_CHECK(res != nullptr, "[KVCache] null result handle");
    res->ctx = kv_cache_create(step, false);
  });
}
int omx_kv_cache_free(omx_kv_cache c) {
  return guarded([&] { kv_cache_destroy((KVCacheImpl*)c.ctx); });
}
int omx_kv_cache_offset(omx_kv_cache c, int* offset) {
  return guarded([&] {
    OMX_CHECK(c.ctx && offset, "[KVCache] null handle");
    *offset = kv_cache_offset((KVCacheImpl*)c.ctx);
  });
}
int omx_kv_cache_reset(omx_kv_cache c) {
  return guarded([&] {
    OMX_CHECK(c.ctx, "[KVCache] null handle");
    kv_cache_reset((KVCacheImpl*)c.ctx);
  });
}
int omx_kv_cache_update_and_fetch(omx_kv_cache c, const omx_array* keys, const omx_array* values,
                                  omx_array* keys_out, omx_array* values_out, omx_stream s) {
  return guarded([&] {
    require_device();
    OMX_CHECK(c.ctx, "[KVCache] null handle");
    OMX_CHECK(!kv_cache_is_concat((KVCacheImpl*)c.ctx), "[KVCache] handle is a ConcatKeyValueCache");
    kv_cache_update((KVCacheImpl*)c.ctx, keys, values, keys_out, values_out, false, (cudaStream_t)s);
  });
}
int omx_kv_cache_state(omx_kv_cache c, omx_array* keys_buf, omx_array* values_buf) {
  return guarded([&] {
    OMX_CHECK(c.ctx, "[KVCache] null handle");
    kv_cache_state((KVCacheImpl*)c.ctx, keys_buf, values_buf);
  });
}
int omx_kv_cache_trim(omx_kv_cache c, int n, int* trimmed) {
  return guarded([&] {
    OMX_CHECK(c.ctx, "[KVCache] null handle");
    const int t = kv_cache_trim((KVCacheImpl*)c.ctx, n);
    if (trimmed) *trimmed = t;
  });
}
int omx_kv_cache_reserve(omx_kv_cache c, int rows) {
  return guarded([&] {
    OMX_CHECK(c.ctx, "[KVCache] null handle");
    kv_cache_reserve((KVCacheImpl*)c.ctx, rows);
  });
}

int omx_concat_kv_cache_new(omx_kv_cache* res) {
  return guarded([&] {
    OMX_CHECK(res != nullptr, "[ConcatKeyValueCache] null result handle");
    res->ctx = kv_cache_create(0, true);
  });
}
int omx_concat_kv_cache_free(omx_kv_cache c) { return omx_kv_cache_free(c); }
int omx_concat_kv_cache_offset(omx_kv_cache c, int* offset) { return omx_kv_cache_offset(c, offset); }
int omx_concat_kv_cache_update_and_fetch(omx_kv_cache c, const omx_array* keys, const omx_array* values,
                                         omx_array* keys_out, omx_array* values_out, omx_stream s) {
  return guarded([&] {
    require_device();
    OMX_CHECK(c.ctx, "[ConcatKeyValueCache] null handle");
    OMX_CHECK(kv_cache_is_concat((KVCacheImpl*)c.ctx), "[ConcatKeyValueCache] handle is a KVCache");
    kv_cache_update((KVCacheImpl*)c.ctx, keys, values, keys_out, values_out, false, (cudaStream_t)s);
  });
}

namespace {
// Shared body of the fused decode step; `peers` != null selects the head-sharded output.
void decode_fused_impl(const omx_array* out, const omx_array* q, const omx_array* k_new,
                       const omx_array* v_new, omx_kv_cache cache, int rope_dims, bool traditional,
                       omx_optional_float base, float rope_scale, const omx_array* freqs, float sm_scale,
                       omx_array* keys_out, omx_array* values_out, const omx_peer_group* peers,
                       int head_offset, cudaStream_t stream, const omx_array* q_norm_w = nullptr,
                       const omx_array* k_norm_w = nullptr, float norm_eps = 0.f, bool peer_wait = false,
                       const omx_ll_group* ll = nullptr) {
  require_device();
  auto* c = (KVCacheImpl*)cache.ctx;
  OMX_CHECK(c, "[attn_decode_fused] null cache handle");
  OMX_CHECK(q && k_new && v_new && out, "[attn_decode_fused] null array");
  OMX_CHECK(q->ndim == 4 && k_new->ndim == 4 && v_new->ndim == 4 && out->ndim == 4,
            "[attn_decode_fused] q, k_new, v_new, out must be [B, H, 1, D]");
  OMX_CHECK(q->shape[2] == 1 && k_new->shape[2] == 1 && v_new->shape[2] == 1,
            "[attn_decode_fused] the fused step handles exactly one new token (L == 1), got L = %lld",
            (long long)q->shape[2]);
  OMX_CHECK(k_new->dtype == q->dtype && v_new->dtype == q->dtype, "[attn_decode_fused] dtype mismatch");
  const int D = (int)q->shape[3];
  OMX_CHECK(rope_dims >= 0 && rope_dims % 2 == 0 && rope_dims <= D, "[rope] dims must be even and <= %d", D);
  OMX_CHECK(rope_dims == 0 || base.has_value != (freqs && freqs->data),
            "[rope] Only one of base or freqs can have a value.");
  const bool qn = q_norm_w && q_norm_w->data, kn = k_norm_w && k_norm_w->data;
  for (const omx_array* w : {qn ? q_norm_w : nullptr, kn ? k_norm_w : nullptr}) {
    if (!w) continue;
    OMX_CHECK(w->ndim == 1 && w->shape[0] == D && w->dtype == q->dtype && (w->strides[0] == 1 || D == 1),
              "[attn_decode_fused_norm] norm weights must be contiguous [%d] vectors in the q dtype", D);
  }
  omx_array out_local = *out;  // the rows of `out` this call writes
  if (peers) {
    OMX_CHECK(peers->world >= 1 && peers->world <= OMX_MAX_PEERS && peers->rank >= 0 && peers->rank < peers->world,
              "[attn_decode_fused_sharded] bad peer group (world %d, rank %d)", peers->world, peers->rank);
    for (int r = 0; r < peers->world; ++r)
      OMX_CHECK(peers->out[r] && peers->flags[r], "[attn_decode_fused_sharded] peer %d is not mapped", r);
    OMX_CHECK(head_offset >= 0 && head_offset + q->shape[1] <= out->shape[1],
              "[attn_decode_fused_sharded] heads [%d, %lld) do not fit the full output (%lld heads)", head_offset,
              (long long)(head_offset + q->shape[1]), (long long)out->shape[1]);
    OMX_CHECK(peers->out[peers->rank] == out->data,
              "[attn_decode_fused_sharded] out_full must be this rank's buffer of the peer group");
    out_local.shape[1] = q->shape[1];
    out_local.data = (char*)out->data + (size_t)head_offset * out->strides[1] * dtype_size(out->dtype);
  }
  if (ll) {
    OMX_CHECK(ll->world >= 1 && ll->world <= OMX_MAX_PEERS && ll->rank >= 0 && ll->rank < ll->world && ll->seq,
              "[attn_decode_fused_sharded_ll] bad group (world %d, rank %d)", ll->world, ll->rank);
    for (int r = 0; r < ll->world; ++r)
      OMX_CHECK(ll->staging[r], "[attn_decode_fused_sharded_ll] staging buffer of rank %d is not mapped", r);
    OMX_CHECK(head_offset == ll->rank * q->shape[1] && (int64_t)ll->world * q->shape[1] == out->shape[1],
              "[attn_decode_fused_sharded_ll] rank %d of %d with %lld local heads must write heads [%lld, ...) of %lld",
              ll->rank, ll->world, (long long)q->shape[1], (long long)(ll->rank * q->shape[1]),
              (long long)out->shape[1]);
    out_local.shape[1] = q->shape[1];
    out_local.data = (char*)out->data + (size_t)head_offset * out->strides[1] * dtype_size(out->dtype);
  }
  const int position = kv_cache_offset(c);
  // cache bookkeeping (growth by the reference rule) without copying the new rows: the kernel
  // ropes k_new and writes row `position` itself.
  omx_array kview, vview;
  KVCacheTxn txn(c, stream);  // a failure below restores offset / capacity
  kv_cache_update(c, k_new, v_new, &kview, &vview, /*skip_copy=*/true, stream);
  SdpaArgs a = make_sdpa_args(&out_local, q, &kview, &vview, sm_scale, "", nullptr, nullptr);
  const char* why = nullptr;
  const bool fast = out->dtype == q->dtype && decode_supported(a, &why) && k_new->strides[3] == 1 &&
                    v_new->strides[3] == 1 && t_forced_kernel != "sdpa_generic";
  // Grouped-query / narrow heads outside head dim 128: ONE prologue launch (norm + rope + row stores) and the
  // mma.sync key-group attention beat the one CUDA-core launch 2 - 4x at batch sizes that matter and are level at one
  // sequence (profiles/r02_mma.md section 4).  Taken only when the prologue kernel covers the layout -- the
  // dynamic-position step decides the same way, so eager and graph replay agree bit for bit; sharded steps stay fused.
  if (fast && t_forced_kernel.empty() && !peers && !ll && out->dtype == q->dtype && v_new->shape[3] == D &&
      !(freqs && freqs->data) && (rope_dims == 0 || base.has_value) && sdpa_mma_preferred_for_decode(a)) {
    const size_t es0 = dtype_size(q->dtype);
    const size_t qb0 = ((size_t)q->shape[0] * q->shape[1] * D * es0 + 255) & ~(size_t)255;
    PrologueCall pc;
    pc.dims = rope_dims;
    pc.traditional = traditional;
    pc.mode = 1;
    pc.eps = norm_eps;
    if (rope_dims > 0) pc.table = get_rope_table(rope_dims, true, base.value, rope_scale, nullptr, position + 1, stream);
    omx_array qt0 = *q;  // contiguous [B, H, 1, D] scratch rows
    qt0.data = get_outer_workspace(qb0, stream);
    qt0.strides[0] = q->shape[1] * D;
    qt0.strides[1] = D;
    qt0.strides[2] = D;
    qt0.strides[3] = 1;
    omx_array krow0 = kview, vrow0 = vview;
    krow0.shape[2] = 1;
    vrow0.shape[2] = 1;
    krow0.data = (char*)kview.data + (size_t)position * kview.strides[2] * dtype_size(kview.dtype);
    vrow0.data = (char*)vview.data + (size_t)position * vview.strides[2] * dtype_size(vview.dtype);
    int n = 0;
    pc.seg[n].x = q; pc.seg[n].out = qt0; pc.seg[n].w = qn ? q_norm_w : nullptr;
    pc.seg[n].rope = rope_dims > 0; pc.seg[n].tok0 = position; ++n;
    pc.seg[n].x = k_new; pc.seg[n].out = krow0; pc.seg[n].w = kn ? k_norm_w : nullptr;
    pc.seg[n].rope = rope_dims > 0; pc.seg[n].tok0 = position; ++n;
    pc.seg[n].x = v_new; pc.seg[n].out = vrow0; pc.seg[n].w = nullptr; pc.seg[n].rope = false; ++n;
    pc.nseg = n;
    note_launch("qkv_prologue");
    if (qkv_prologue(pc, stream)) {
      SdpaArgs a2 = make_sdpa_args(&out_local, &qt0, &kview, &vview, sm_scale, "", nullptr, nullptr);
      sdpa_mma(a2, stream);
      txn.commit();
      if (keys_out) *keys_out = kview;
      if (values_out) *values_out = vview;
      return;
    }
  }
  OMX_CHECK(fast || (!peers && !ll), "[attn_decode_fused_sharded] layout not supported by the decode kernels: %s",
            why ? why : "strided k_new/v_new");
  if (fast) {
    DecodeFused f;
    f.enabled = true;
    f.k_new = k_new;
    f.v_new = v_new;
    f.rope_dims = rope_dims;
    f.traditional = traditional;
    f.position = position;
    f.stable_rows = kv_cache_stable_rows(c);
    f.peers = peers;
    f.peer_wait = peer_wait;
    f.ll = ll;
    f.ll_out_full = out->data;
    f.head_offset = 0;  // out_local already starts at this rank's first head
    f.q_norm_w = qn ? q_norm_w->data : nullptr;
    f.k_norm_w = kn ? k_norm_w->data : nullptr;
    f.norm_eps = norm_eps;
    if (rope_dims > 0) {
      std::vector<float> fh;
      if (!base.has_value) {
        OMX_CHECK(freqs->ndim == 1 && freqs->shape[0] == rope_dims / 2 && freqs->dtype == OMX_FLOAT32 &&
                      freqs->strides[0] == 1,
                  "[rope] freqs must be a contiguous float32 vector of length dims/2");
        fh.resize(rope_dims / 2);
        OMX_CUDA(cudaMemcpyAsync(fh.data(), freqs->data, sizeof(float) * fh.size(), cudaMemcpyDeviceToHost,
                                 stream));
        OMX_CUDA(cudaStreamSynchronize(stream));
      }
      f.table = get_rope_table(rope_dims, base.has_value, base.value, rope_scale,
                               fh.empty() ? nullptr : fh.data(), position + 1, stream);
    }
    omx_peer_group shifted;
    if (peers) {  // peer pointers address the same head slice in every rank's buffer
      shifted = *peers;
      const size_t shift = (size_t)head_offset * out->strides[1] * dtype_size(out->dtype);
      for (int r = 0; r < peers->world; ++r) shifted.out[r] = (char*)peers->out[r] + shift;
      f.peers = &shifted;
    }
    decode_attention(a, f, stream);
    txn.commit();
    if (keys_out) *keys_out = kview;
    if (values_out) *values_out = vview;
    return;
  }
  // Unfused composition for layouts the decode kernels do not take: (norm) -> rope -> row store -> sdpa.
  const size_t es = dtype_size(q->dtype);
  const size_t qbytes = ((size_t)q->shape[0] * q->shape[1] * D * es + 255) & ~(size_t)255;
  const size_t kbytes = ((size_t)k_new->shape[0] * k_new->shape[1] * D * es + 255) & ~(size_t)255;
  char* ws = (qn || kn || rope_dims > 0) ? (char*)get_outer_workspace(2 * qbytes + kbytes, stream) : nullptr;
  auto dense = [&](const omx_array* like, char* mem) {  // contiguous [B,H,1,D] scratch array
    omx_array t = *like;
    t.data = mem;
    t.strides[0] = like->shape[1] * D;
    t.strides[1] = D;
    t.strides[2] = D;
    t.strides[3] = 1;
    return t;
  };
  auto sdpa_rows = [&](const SdpaArgs& a2) {  // 16-bit rows on tensor cores where the layout allows, else one warp per row
    const char* why2 = nullptr;
    if (t_forced_kernel != "sdpa_generic" && sdpa_mma_supported(a2, &why2)) sdpa_mma(a2, stream);
    else sdpa_generic(a2, stream);
  };
  {
    // One launch for norm + rope + the row stores when the layout allows (prologue.cu), like the prefill composite:
    // k' -> cache row, v -> cache row, q' -> scratch; then the attention.  Otherwise the standalone ops below.
    omx_array krow0 = kview, vrow0 = vview;
    krow0.shape[2] = 1;
    vrow0.shape[2] = 1;
    krow0.data = (char*)kview.data + (size_t)position * kview.strides[2] * dtype_size(kview.dtype);
    vrow0.data = (char*)vview.data + (size_t)position * vview.strides[2] * dtype_size(vview.dtype);
    const bool has_freqs = freqs && freqs->data;
    if (!has_freqs && (rope_dims == 0 || base.has_value) && t_forced_kernel != "sdpa_generic") {
      PrologueCall pc;
      pc.dims = rope_dims;
      pc.traditional = traditional;
      pc.mode = 1;
      pc.eps = norm_eps;
      if (rope_dims > 0)
        pc.table = get_rope_table(rope_dims, true, base.value, rope_scale, nullptr, position + 1, stream);
      int n = 0;
      omx_array qt0{};
      if (qn || rope_dims > 0) {
        qt0 = dense(q, ws + qbytes);
        pc.seg[n].x = q; pc.seg[n].out = qt0; pc.seg[n].w = qn ? q_norm_w : nullptr;
        pc.seg[n].rope = rope_dims > 0; pc.seg[n].tok0 = position; ++n;
      }
      pc.seg[n].x = k_new; pc.seg[n].out = krow0; pc.seg[n].w = kn ? k_norm_w : nullptr;
      pc.seg[n].rope = rope_dims > 0; pc.seg[n].tok0 = position; ++n;
      const bool v_in = v_new->shape[3] == D;
      if (v_in) {
        pc.seg[n].x = v_new; pc.seg[n].out = vrow0; pc.seg[n].w = nullptr; pc.seg[n].rope = false; ++n;
      }
      pc.nseg = n;
      note_launch("qkv_prologue");
      if (qkv_prologue(pc, stream)) {
        if (!v_in) copy4d(&vrow0, v_new, stream);
        SdpaArgs a2 = make_sdpa_args(out, (qn || rope_dims > 0) ? &qt0 : q, &kview, &vview, sm_scale, "", nullptr, nullptr);
        sdpa_rows(a2);
        txn.commit();
        if (keys_out) *keys_out = kview;
        if (values_out) *values_out = vview;
        return;
      }
    }
  }
  omx_array qn_arr, kn_arr;
  if (qn) {
    qn_arr = dense(q, ws);
    rms_norm_forward(&qn_arr, q, q_norm_w, norm_eps, stream);
    q = &qn_arr;
  }
  if (kn) {
    kn_arr = dense(k_new, ws + 2 * qbytes);
    rms_norm_forward(&kn_arr, k_new, k_norm_w, norm_eps, stream);
    k_new = &kn_arr;
  }
  omx_array krow = kview, vrow = vview;
  krow.shape[2] = 1;
  vrow.shape[2] = 1;
  krow.data = (char*)kview.data + (size_t)position * kview.strides[2] * dtype_size(kview.dtype);
  vrow.data = (char*)vview.data + (size_t)position * vview.strides[2] * dtype_size(vview.dtype);
  if (rope_dims > 0) {
    rope_forward(&krow, k_new, rope_dims, traditional, base, rope_scale, position, nullptr, 0, freqs, stream);
  } else {
    copy4d(&krow, k_new, stream);
  }
  copy4d(&vrow, v_new, stream);
  if (rope_dims > 0) {
    omx_array qr = dense(q, ws + qbytes);
    rope_forward(&qr, q, rope_dims, traditional, base, rope_scale, position, nullptr, 0, freqs, stream);
    SdpaArgs a2 = make_sdpa_args(out, &qr, &kview, &vview, sm_scale, "", nullptr, nullptr);
    sdpa_rows(a2);
  } else {
    SdpaArgs a2 = make_sdpa_args(out, q, &kview, &vview, sm_scale, "", nullptr, nullptr);
    sdpa_rows(a2);
  }
  txn.commit();
  if (keys_out) *keys_out = kview;
  if (values_out) *values_out = vview;
}
}  // namespace

int omx_attn_decode_fused(const omx_array* out, const omx_array* q, const omx_array* k_new,
                          const omx_array* v_new, omx_kv_cache cache, int rope_dims, bool traditional,
                          omx_optional_float base, float rope_scale, const omx_array* freqs, float sm_scale,
                          omx_array* keys_out, omx_array* values_out, omx_stream s) {
  return guarded([&] {
    decode_fused_impl(out, q, k_new, v_new, cache, rope_dims, traditional, base, rope_scale, freqs, sm_scale,
                      keys_out, values_out, nullptr, 0, (cudaStream_t)s);
  });
}

int omx_attn_decode_fused_norm(const omx_array* out, const omx_array* q, const omx_array* k_new,
                               const omx_array* v_new, omx_kv_cache cache, const omx_array* q_norm_weight,
                               const omx_array* k_norm_weight, float norm_eps, int rope_dims, bool traditional,
                               omx_optional_float base, float rope_scale, const omx_array* freqs, float sm_scale,
                               omx_array* keys_out, omx_array* values_out, omx_stream s) {
  return guarded([&] {
    decode_fused_impl(out, q, k_new, v_new, cache, rope_dims, traditional, base, rope_scale, freqs, sm_scale,
                      keys_out, values_out, nullptr, 0, (cudaStream_t)s, q_norm_weight, k_norm_weight, norm_eps);
  });
}

static __global__ void omx_counter_add_kernel(int32_t* c, int delta) { *c += delta; }

int omx_device_counter_add(int32_t* counter, int delta, omx_stream s) {
  return guarded([&] {
    require_device();
    OMX_CHECK(counter, "[device_counter_add] null counter");
    omx_counter_add_kernel<<<1, 1, 0, (cudaStream_t)s>>>(counter, delta);
    count_launch();
    OMX_CUDA(cudaGetLastError());
  });
}

int omx_kv_cache_prepare_graph(omx_kv_cache c, int max_rows, int n_q_heads, omx_stream s) {
  return guarded([&] {
    require_device();
    auto* kc = (KVCacheImpl*)c.ctx;
    OMX_CHECK(kc, "[KVCache] null handle");
    int B, H, Dk, Dv, dt;
    kv_cache_shape(kc, &B, &H, &Dk, &Dv, &dt);
    OMX_CHECK(n_q_heads >= H && n_q_heads % H == 0, "[KVCache] prepare_graph: %d query heads over %d kv heads",
              n_q_heads, H);
    // scratch for whichever kernels serve the dynamic-position step: the decode kernels' split-K partials, or the
    // prologue's q' rows + the mma.sync kernel's partials
    const size_t qbytes = ((size_t)B * n_q_heads * Dk * dtype_size(dt) + 255) & ~(size_t)255;
    // (only geometries that can take the mma.sync route pay for its scratch: 16-bit, not the 128 / 128 heads of the TMA kernel)
    const bool mma_route = (dt == OMX_BFLOAT16 || dt == OMX_FLOAT16) && !(Dk == 128 && Dv == 128);
    const size_t need = std::max(decode_graph_scratch_bytes(B, H, n_q_heads, Dk, dt, max_rows),
                                 mma_route ? qbytes + sdpa_mma_graph_scratch_bytes(B, H, n_q_heads, Dv) : (size_t)0);
    kv_cache_prepare_graph(kc, max_rows, need, (cudaStream_t)s);
  });
}

int omx_kv_cache_advance(omx_kv_cache c, int n, omx_stream s) {
  return guarded([&] {
    OMX_CHECK(c.ctx, "[KVCache] null handle");
    kv_cache_advance((KVCacheImpl*)c.ctx, n, (cudaStream_t)s);
  });
}

int omx_attn_decode_fused_dynamic(const omx_array* out, const omx_array* q, const omx_array* k_new,
                                  const omx_array* v_new, omx_kv_cache cache, const omx_array* q_norm_weight,
                                  const omx_array* k_norm_weight, float norm_eps, int rope_dims, bool traditional,
                                  omx_optional_float base, float rope_scale, float sm_scale,
                                  const int32_t* position, omx_stream s) {
  return guarded([&] {
    require_device();
    cudaStream_t stream = (cudaStream_t)s;
    auto* c = (KVCacheImpl*)cache.ctx;
    OMX_CHECK(c, "[attn_decode_fused_dynamic] null cache handle");
    OMX_CHECK(position, "[attn_decode_fused_dynamic] null position pointer");
    OMX_CHECK(q && k_new && v_new && out, "[attn_decode_fused_dynamic] null array");
    OMX_CHECK(q->ndim == 4 && k_new->ndim == 4 && v_new->ndim == 4 && out->ndim == 4 && q->shape[2] == 1 &&
                  k_new->shape[2] == 1 && v_new->shape[2] == 1,
              "[attn_decode_fused_dynamic] q, k_new, v_new, out must be [B, H, 1, D]");
    OMX_CHECK(k_new->dtype == q->dtype && v_new->dtype == q->dtype && out->dtype == q->dtype,
              "[attn_decode_fused_dynamic] dtype mismatch");
    const int D = (int)q->shape[3];
    OMX_CHECK(rope_dims >= 0 && rope_dims % 2 == 0 && rope_dims <= D, "[rope] dims must be even and <= %d", D);
    OMX_CHECK(rope_dims == 0 || base.has_value, "[attn_decode_fused_dynamic] rope needs a base (no freqs here)");
    const bool qn = q_norm_weight && q_norm_weight->data, kn = k_norm_weight && k_norm_weight->data;
    for (const omx_array* w : {qn ? q_norm_weight : nullptr, kn ? k_norm_weight : nullptr}) {
      if (!w) continue;
      OMX_CHECK(w->ndim == 1 && w->shape[0] == D && w->dtype == q->dtype && (w->strides[0] == 1 || D == 1),
                "[attn_decode_fused_dynamic] norm weights must be contiguous [%d] vectors in the q dtype", D);
    }
    omx_array kview, vview;
    DecodeFused f;
    OMX_CHECK(kv_cache_graph_view(c, &kview, &vview, &f.scratch, &f.scratch_bytes, &f.max_rows),
              "[attn_decode_fused_dynamic] call omx_kv_cache_prepare_graph first (or again: the cache grew past "
              "the pinned rows)");
    OMX_CHECK(k_new->shape[0] == kview.shape[0] && k_new->shape[1] == kview.shape[1] && k_new->shape[3] == kview.shape[3] &&
                  v_new->shape[3] == vview.shape[3] && k_new->dtype == kview.dtype,
              "[attn_decode_fused_dynamic] k_new / v_new do not match the cache shape or dtype");
    SdpaArgs a = make_sdpa_args(out, q, &kview, &vview, sm_scale, "", nullptr, nullptr);
    const char* why = nullptr;
    const bool fast = decode_supported(a, &why) && k_new->strides[3] == 1 && v_new->strides[3] == 1;
    OMX_CHECK(fast, "[attn_decode_fused_dynamic] layout not supported by the decode kernels: %s",
              why ? why : "strided k_new/v_new");
    // Grouped-query / narrow heads outside head dim 128: prologue + mma.sync key groups, as the eager step does
    // (same kernels, same split plan derived from the device position: bit-identical outputs and cache rows)
    if (t_forced_kernel.empty() && sdpa_mma_preferred_for_decode(a) && v_new->shape[3] == D) {
      const size_t es = dtype_size(q->dtype);
      const size_t qbytes = ((size_t)q->shape[0] * q->shape[1] * D * es + 255) & ~(size_t)255;
      if (f.scratch && f.scratch_bytes > qbytes) {
        PrologueCall pc;
        pc.dims = rope_dims;
        pc.traditional = traditional;
        pc.mode = 1;
        pc.eps = norm_eps;
        pc.pos_dev = position;
        if (rope_dims > 0) pc.table = get_rope_table(rope_dims, true, base.value, rope_scale, nullptr, f.max_rows, stream);
        omx_array qt0 = *q;  // contiguous [B, H, 1, D] rows at the head of the scratch
        qt0.data = f.scratch;
        qt0.strides[0] = q->shape[1] * D;
        qt0.strides[1] = D;
        qt0.strides[2] = D;
        qt0.strides[3] = 1;
        omx_array krow0 = kview, vrow0 = vview;  // row 0 of the pinned views; the kernel moves to row *position
        krow0.shape[2] = 1;
        vrow0.shape[2] = 1;
        int n = 0;
        pc.seg[n].x = q; pc.seg[n].out = qt0; pc.seg[n].w = qn ? q_norm_weight : nullptr;
        pc.seg[n].rope = rope_dims > 0; pc.seg[n].tok0 = 0; ++n;
        pc.seg[n].x = k_new; pc.seg[n].out = krow0; pc.seg[n].w = kn ? k_norm_weight : nullptr;
        pc.seg[n].rope = rope_dims > 0; pc.seg[n].tok0 = 0; pc.seg[n].dyn_row_stride = kview.strides[2]; ++n;
        pc.seg[n].x = v_new; pc.seg[n].out = vrow0; pc.seg[n].w = nullptr; pc.seg[n].rope = false;
        pc.seg[n].dyn_row_stride = vview.strides[2]; ++n;
        pc.nseg = n;
        note_launch("qkv_prologue");
        if (qkv_prologue(pc, stream)) {
          SdpaArgs a2 = make_sdpa_args(out, &qt0, &kview, &vview, sm_scale, "", nullptr, nullptr);
          sdpa_mma_dynamic(a2, position, (char*)f.scratch + qbytes, f.scratch_bytes - qbytes, stream);
          return;
        }
      }
    }
    f.enabled = true;
    f.k_new = k_new;
    f.v_new = v_new;
    f.rope_dims = rope_dims;
    f.traditional = traditional;
    f.position = 0;  // table base; the kernel adds *position rows
    f.pos_dev = position;
    f.q_norm_w = qn ? q_norm_weight->data : nullptr;
    f.k_norm_w = kn ? k_norm_weight->data : nullptr;
    f.norm_eps = norm_eps;
    if (rope_dims > 0) f.table = get_rope_table(rope_dims, true, base.value, rope_scale, nullptr, f.max_rows, stream);
    decode_attention(a, f, stream);
  });
}

// ---- paged KV cache (paged_kv.cu) ----
int omx_paged_kv_cache_new(omx_paged_kv_cache* res, int batch, int n_kv_heads, int head_dim_k, int head_dim_v, int dtype,
                           int64_t n_pages, int max_pages_per_seq) {
  return guarded([&] {
    require_device();
    OMX_CHECK(res != nullptr, "[PagedKVCache] null result handle");
    res->ctx = paged_create(batch, n_kv_heads, head_dim_k, head_dim_v, dtype, n_pages, max_pages_per_seq);
  });
}
int omx_paged_kv_cache_free(omx_paged_kv_cache c) {
  return guarded([&] { paged_destroy((PagedKVImpl*)c.ctx); });
}
#define OMX_PAGED(c) \
  auto* pc = (PagedKVImpl*)(c).ctx; \
  OMX_CHECK(pc, "[PagedKVCache] null handle")
int omx_paged_kv_cache_offset(omx_paged_kv_cache c, int* offset) {
  return guarded([&] {
    OMX_PAGED(c);
    OMX_CHECK(offset, "[PagedKVCache] null result pointer");
    *offset = std::max(paged_offset(pc), 0);
  });
}
int omx_paged_kv_cache_lengths(omx_paged_kv_cache c, int32_t* lens) {
  return guarded([&] {
    OMX_PAGED(c);
    OMX_CHECK(lens, "[PagedKVCache] null result pointer");
    std::copy(paged_lengths(pc), paged_lengths(pc) + paged_batch(pc), lens);
  });
}
int omx_paged_kv_cache_free_pages(omx_paged_kv_cache c, int64_t* n) {
  return guarded([&] {
    OMX_PAGED(c);
    OMX_CHECK(n, "[PagedKVCache] null result pointer");
    *n = paged_free_pages(pc);
  });
}
int omx_paged_kv_cache_reset(omx_paged_kv_cache c, int slot, omx_stream s) {
  return guarded([&] {
    require_device();
    OMX_PAGED(c);
    paged_reset(pc, slot, false, (cudaStream_t)s);
  });
}
int omx_paged_kv_cache_release(omx_paged_kv_cache c, int slot, omx_stream s) {
  return guarded([&] {
    require_device();
    OMX_PAGED(c);
    OMX_CHECK(slot >= 0, "[PagedKVCache] release needs a slot index");
    paged_reset(pc, slot, true, (cudaStream_t)s);
  });
}
int omx_paged_kv_cache_reserve(omx_paged_kv_cache c, int rows_ahead, omx_stream s) {
  return guarded([&] {
    require_device();
    OMX_PAGED(c);
    paged_reserve(pc, rows_ahead, (cudaStream_t)s);
  });
}
int omx_paged_kv_cache_sync_lengths(omx_paged_kv_cache c, omx_stream s) {
  return guarded([&] {
    require_device();
    OMX_PAGED(c);
    paged_sync_lengths(pc, (cudaStream_t)s);
  });
}
int omx_paged_kv_cache_trim(omx_paged_kv_cache c, int n, omx_stream s) {
  return guarded([&] {
    require_device();
    OMX_PAGED(c);
    paged_trim(pc, n, (cudaStream_t)s);
  });
}
int omx_paged_kv_cache_update_and_fetch(omx_paged_kv_cache c, const omx_array* keys, const omx_array* values,
                                        omx_array* keys_out, omx_array* values_out, omx_stream s) {
  return guarded([&] {
    require_device();
    OMX_PAGED(c);
    OMX_CHECK(keys && keys->ndim == 4 && keys->shape[0] == paged_batch(pc),
              "[PagedKVCache] update_and_fetch appends to every sequence: keys must be [%d, n_kv_heads, n, head_dim]",
              paged_batch(pc));
    paged_append(pc, 0, keys, values, (cudaStream_t)s);
    if (keys_out || values_out) paged_materialize(pc, keys_out, values_out, (cudaStream_t)s);
  });
}
int omx_paged_kv_cache_append_slot(omx_paged_kv_cache c, int slot, const omx_array* keys, const omx_array* values,
                                   omx_stream s) {
  return guarded([&] {
    require_device();
    OMX_PAGED(c);
    OMX_CHECK(keys && keys->ndim == 4 && keys->shape[0] == 1, "[PagedKVCache] append_slot takes [1, n_kv_heads, n, head_dim]");
    paged_append(pc, slot, keys, values, (cudaStream_t)s);
  });
}
int omx_paged_kv_cache_fetch(omx_paged_kv_cache c, omx_array* keys_out, omx_array* values_out, omx_stream s) {
  return guarded([&] {
    require_device();
    OMX_PAGED(c);
    paged_materialize(pc, keys_out, values_out, (cudaStream_t)s);
  });
}
int omx_paged_kv_cache_pages(omx_paged_kv_cache c, void** k_pool, void** v_pool, const int32_t** block_table,
                             int* max_pages_per_seq) {
  return guarded([&] {
    OMX_PAGED(c);
    void *kp, *vp;
    const int* bt;
    paged_pool_ptrs(pc, &kp, &vp, &bt);
    if (k_pool) *k_pool = kp;
    if (v_pool) *v_pool = vp;
    if (block_table) *block_table = bt;
    int B, H, Dk, Dv, dt, mp;
    int64_t np;
    paged_shape(pc, &B, &H, &Dk, &Dv, &dt, &np, &mp);
    if (max_pages_per_seq) *max_pages_per_seq = mp;
  });
}

int omx_attn_decode_fused_paged(const omx_array* out, const omx_array* q, const omx_array* k_new,
                                const omx_array* v_new, omx_paged_kv_cache cache, const omx_array* q_norm_weight,
                                const omx_array* k_norm_weight, float norm_eps, int rope_dims, bool traditional,
                                omx_optional_float base, float rope_scale, float sm_scale, omx_stream s) {
  return guarded([&] {
    require_device();
    cudaStream_t stream = (cudaStream_t)s;
    OMX_PAGED(cache);
    int B, H, Dk, Dv, dt, max_pages;
    int64_t n_pages;
    paged_shape(pc, &B, &H, &Dk, &Dv, &dt, &n_pages, &max_pages);
    OMX_CHECK(q && k_new && v_new && out, "[attn_decode_fused_paged] null array");
    OMX_CHECK(q->ndim == 4 && k_new->ndim == 4 && v_new->ndim == 4 && out->ndim == 4 && q->shape[2] == 1 &&
                  k_new->shape[2] == 1 && v_new->shape[2] == 1,
              "[attn_decode_fused_paged] q, k_new, v_new, out must be [B, H, 1, D]");
    OMX_CHECK(q->dtype == dt && k_new->dtype == dt && v_new->dtype == dt && out->dtype == dt,
              "[attn_decode_fused_paged] q, k_new, v_new, out must have the cache dtype");
    OMX_CHECK(q->shape[0] == B && k_new->shape[0] == B && v_new->shape[0] == B && k_new->shape[1] == H &&
                  v_new->shape[1] == H && k_new->shape[3] == Dk && v_new->shape[3] == Dv && q->shape[3] == Dk,
              "[attn_decode_fused_paged] arrays do not match the cache geometry [%d, %d, *, %d]", B, H, Dk);
    const int D = Dk;
    OMX_CHECK(rope_dims >= 0 && rope_dims % 2 == 0 && rope_dims <= D, "[rope] dims must be even and <= %d", D);
    OMX_CHECK(rope_dims == 0 || base.has_value, "[attn_decode_fused_paged] rope needs a base (no freqs here)");
    const bool qn = q_norm_weight && q_norm_weight->data, kn = k_norm_weight && k_norm_weight->data;
    for (const omx_array* w : {qn ? q_norm_weight : nullptr, kn ? k_norm_weight : nullptr}) {
      if (!w) continue;
      OMX_CHECK(w->ndim == 1 && w->shape[0] == D && w->dtype == q->dtype && (w->strides[0] == 1 || D == 1),
                "[attn_decode_fused_paged] norm weights must be contiguous [%d] vectors in the q dtype", D);
    }
    OMX_CHECK(q->shape[1] % H == 0, "[scaled_dot_product_attention] n_heads must be a multiple of n_kv_heads, found "
              "n_heads %lld for n_kv_heads %d", (long long)q->shape[1], H);
    OMX_CHECK(out->shape[0] == B && out->shape[1] == q->shape[1] && out->shape[2] == 1 && out->shape[3] == Dv,
              "[scaled_dot_product_attention] out must be [B, n_heads, L_q, D_v]");
    PagedRef ref;
    DecodeFused f;
    omx_array kview, vview;
    int max_len_after = 0, table_rows = 0;
    paged_begin_step(pc, (int)q->shape[1], &kview, &vview, &ref, &f.scratch, &f.scratch_bytes, &max_len_after,
                     &table_rows, stream);
    if (max_len_after == 0) return;  // every slot released: nothing to do
    SdpaArgs a = make_sdpa_args(out, q, &kview, &vview, sm_scale, "", nullptr, nullptr);
    const char* why = nullptr;
    const bool fast = decode_supported(a, &why) && k_new->strides[3] == 1 && v_new->strides[3] == 1;
    OMX_CHECK(fast, "[attn_decode_fused_paged] layout not supported by the decode kernels: %s",
              why ? why : "strided k_new/v_new");
    // Grouped-query / narrow heads outside head dim 128: prologue + mma.sync key groups, like the contiguous cache's
    // step (same kernels, split plan from each sequence's own length: same bits)
    if (t_forced_kernel.empty() && sdpa_mma_preferred_for_decode(a) && Dv == D) {
      const size_t qbytes = ((size_t)B * q->shape[1] * D * dtype_size(dt) + 255) & ~(size_t)255;
      if (f.scratch && f.scratch_bytes > qbytes) {
        PrologueCall pro;
        pro.dims = rope_dims;
        pro.traditional = traditional;
        pro.mode = 1;
        pro.eps = norm_eps;
        pro.paged = &ref;
        if (rope_dims > 0) pro.table = get_rope_table(rope_dims, true, base.value, rope_scale, nullptr, table_rows, stream);
        omx_array qt0 = *q;  // contiguous [B, H, 1, D] rows at the head of the scratch
        qt0.data = f.scratch;
        qt0.strides[0] = q->shape[1] * D;
        qt0.strides[1] = D;
        qt0.strides[2] = D;
        qt0.strides[3] = 1;
        omx_array krow0 = kview, vrow0 = vview;  // the pools: page / head / row strides; the kernel picks the page row
        krow0.shape[2] = 1;
        vrow0.shape[2] = 1;
        int n = 0;
        pro.seg[n].x = q; pro.seg[n].out = qt0; pro.seg[n].w = qn ? q_norm_weight : nullptr;
        pro.seg[n].rope = rope_dims > 0; pro.seg[n].tok0 = 0; ++n;
        pro.seg[n].x = k_new; pro.seg[n].out = krow0; pro.seg[n].w = kn ? k_norm_weight : nullptr;
        pro.seg[n].rope = rope_dims > 0; pro.seg[n].tok0 = 0; pro.seg[n].paged_dst = true; ++n;
        pro.seg[n].x = v_new; pro.seg[n].out = vrow0; pro.seg[n].w = nullptr; pro.seg[n].rope = false;
        pro.seg[n].paged_dst = true; ++n;
        pro.nseg = n;
        note_launch("qkv_prologue");
        if (qkv_prologue(pro, stream)) {
          SdpaArgs a2 = make_sdpa_args(out, &qt0, &kview, &vview, sm_scale, "", nullptr, nullptr);
          sdpa_mma_dynamic(a2, nullptr, (char*)f.scratch + qbytes, f.scratch_bytes - qbytes, stream, &ref);
          paged_end_step(pc);
          return;
        }
      }
    }
    f.enabled = true;
    f.k_new = k_new;
    f.v_new = v_new;
    f.rope_dims = rope_dims;
    f.traditional = traditional;
    f.position = 0;  // table base; the kernel adds len[b] rows
    f.paged = &ref;
    f.q_norm_w = qn ? q_norm_weight->data : nullptr;
    f.k_norm_w = kn ? k_norm_weight->data : nullptr;
    f.norm_eps = norm_eps;
    if (rope_dims > 0) f.table = get_rope_table(rope_dims, true, base.value, rope_scale, nullptr, table_rows, stream);
    decode_attention(a, f, stream);
    paged_end_step(pc);
  });
}

int omx_attn_prefill_fused(const omx_array* out, const omx_array* q, const omx_array* k_new,
                           const omx_array* v_new, omx_kv_cache cache, const omx_array* q_norm_weight,
                           const omx_array* k_norm_weight, float norm_eps, int rope_dims, bool traditional,
                           omx_optional_float base, float rope_scale, const omx_array* freqs, float sm_scale,
                           const char* mask_mode, const omx_array* mask_arr, omx_array* keys_out,
                           omx_array* values_out, omx_stream s) {
  return guarded([&] {
    require_device();
    cudaStream_t stream = (cudaStream_t)s;
    auto* c = (KVCacheImpl*)cache.ctx;
    OMX_CHECK(c, "[attn_prefill_fused] null cache handle");
    OMX_CHECK(q && k_new && v_new && out, "[attn_prefill_fused] null array");
    OMX_CHECK(q->ndim == 4 && k_new->ndim == 4 && v_new->ndim == 4 && out->ndim == 4,
              "[attn_prefill_fused] q, k_new, v_new, out must be [B, H, L, D]");
    OMX_CHECK(k_new->dtype == q->dtype && v_new->dtype == q->dtype, "[attn_prefill_fused] dtype mismatch");
    OMX_CHECK(q->shape[2] == k_new->shape[2] && q->shape[2] == v_new->shape[2],
              "[attn_prefill_fused] q, k_new, v_new must carry the same number of new tokens");
    const int D = (int)q->shape[3];
    const int64_t L = q->shape[2];
    OMX_CHECK(rope_dims >= 0 && rope_dims % 2 == 0 && rope_dims <= D, "[rope] dims must be even and <= %d", D);
    const bool qn = q_norm_weight && q_norm_weight->data, kn = k_norm_weight && k_norm_weight->data;
    const int position = kv_cache_offset(c);
    // cache bookkeeping by the reference rule; the rows are written below, straight into the cache
    omx_array kview, vview;
    OMX_CHECK(rope_dims == 0 || base.has_value != (freqs && freqs->data),
              "[rope] Only one of base or freqs can have a value.");
    KVCacheTxn txn(c, stream);  // a failure below restores offset / capacity
    kv_cache_update(c, k_new, v_new, &kview, &vview, /*skip_copy=*/true, stream);
    if ((int64_t)q->shape[0] * L == 0) {
      txn.commit();
      if (keys_out) *keys_out = kview;
      if (values_out) *values_out = vview;
      return;
    }
    auto rows = [&](const omx_array& view) {  // rows [position, position + L) of a fetched view
      omx_array r = view;
      r.shape[2] = L;
      r.data = (char*)view.data + (size_t)position * view.strides[2] * dtype_size(view.dtype);
      return r;
    };
    omx_array krows = rows(kview), vrows = rows(vview);
    // validate the attention call (out shape, mask, head counts) BEFORE the prologue writes any cache row
    (void)make_sdpa_args(out, q, &kview, &vview, sm_scale, mask_mode, mask_arr, nullptr);
    const size_t es = dtype_size(q->dtype);
    const size_t qbytes = ((size_t)q->shape[0] * q->shape[1] * L * D * es + 255) & ~(size_t)255;
    const size_t kbytes = ((size_t)k_new->shape[0] * k_new->shape[1] * L * D * es + 255) & ~(size_t)255;
    char* ws = (char*)get_outer_workspace(qbytes + (kn && rope_dims > 0 ? kbytes : 0), stream);
    auto dense = [&](const omx_array* like, char* mem) {
      omx_array t = *like;
      t.data = mem;
      t.strides[3] = 1;
      t.strides[2] = D;
      t.strides[1] = L * D;
      t.strides[0] = like->shape[1] * L * D;
      return t;
    };
    // One launch for the whole prologue when the layout allows (prologue.cu): k' -> cache rows,
    // v -> cache rows, q' -> scratch.  Otherwise the standalone ops below, same results.
    const bool has_freqs = freqs && freqs->data;
    bool fused_prologue = false;
    if (!has_freqs && (rope_dims == 0 || base.has_value)) {
      PrologueCall pc;
      pc.dims = rope_dims;
      pc.traditional = traditional;
      pc.mode = 1;
      pc.eps = norm_eps;
      if (rope_dims > 0)
        pc.table = get_rope_table(rope_dims, true, base.value, rope_scale, nullptr, position + (int)L, stream);
      int n = 0;
      omx_array qt0{};
      if (qn || rope_dims > 0) {
        qt0 = dense(q, ws);
        pc.seg[n].x = q; pc.seg[n].out = qt0; pc.seg[n].w = qn ? q_norm_weight : nullptr;
        pc.seg[n].rope = rope_dims > 0; pc.seg[n].tok0 = position; ++n;
      }
      pc.seg[n].x = k_new; pc.seg[n].out = krows; pc.seg[n].w = kn ? k_norm_weight : nullptr;
      pc.seg[n].rope = rope_dims > 0; pc.seg[n].tok0 = position; ++n;
      const bool v_in = v_new->shape[3] == D;
      if (v_in) {
        pc.seg[n].x = v_new; pc.seg[n].out = vrows; pc.seg[n].w = nullptr; pc.seg[n].rope = false; ++n;
      }
      pc.nseg = n;
      note_launch("qkv_prologue");
      if (qkv_prologue(pc, stream)) {
        fused_prologue = true;
        if (!v_in) copy4d(&vrows, v_new, stream);
        const omx_array* qa = (qn || rope_dims > 0) ? &qt0 : q;
        SdpaArgs a = make_sdpa_args(out, qa, &kview, &vview, sm_scale, mask_mode, mask_arr, nullptr);
        dispatch_sdpa(a, stream);
      }
    }
    if (fused_prologue) {
      txn.commit();
      if (keys_out) *keys_out = kview;
      if (values_out) *values_out = vview;
      return;
    }
    // k: (norm) -> rope -> cache rows; v: copy -> cache rows; q: (norm) -> rope -> scratch
    if (kn && rope_dims > 0) {
      omx_array kt = dense(k_new, ws + qbytes);
      rms_norm_forward(&kt, k_new, k_norm_weight, norm_eps, stream);
      rope_forward(&krows, &kt, rope_dims, traditional, base, rope_scale, position, nullptr, 0, freqs, stream);
    } else if (kn) {
      rms_norm_forward(&krows, k_new, k_norm_weight, norm_eps, stream);
    } else if (rope_dims > 0) {
      rope_forward(&krows, k_new, rope_dims, traditional, base, rope_scale, position, nullptr, 0, freqs, stream);
    } else {
      copy4d(&krows, k_new, stream);
    }
    copy4d(&vrows, v_new, stream);
    const omx_array* qa = q;
    omx_array qt;
    if (qn || rope_dims > 0) {
      qt = dense(q, ws);
      if (qn) rms_norm_forward(&qt, q, q_norm_weight, norm_eps, stream);
      if (rope_dims > 0)
        rope_forward(&qt, qn ? &qt : q, rope_dims, traditional, base, rope_scale, position, nullptr, 0, freqs, stream);
      qa = &qt;
    }
    SdpaArgs a = make_sdpa_args(out, qa, &kview, &vview, sm_scale, mask_mode, mask_arr, nullptr);
    dispatch_sdpa(a, stream);
    txn.commit();
    if (keys_out) *keys_out = kview;
    if (values_out) *values_out = vview;
  });
}

int omx_attn_decode_fused_sharded(const omx_array* out_full, const omx_array* q, const omx_array* k_new,
                                  const omx_array* v_new, omx_kv_cache cache, int rope_dims, bool traditional,
                                  omx_optional_float base, float rope_scale, const omx_array* freqs,
                                  float sm_scale, const omx_peer_group* peers, int head_offset, omx_stream s) {
  return guarded([&] {
    OMX_CHECK(peers != nullptr, "[attn_decode_fused_sharded] null peer group");
    decode_fused_impl(out_full, q, k_new, v_new, cache, rope_dims, traditional, base, rope_scale, freqs,
                      sm_scale, nullptr, nullptr, peers, head_offset, (cudaStream_t)s);
  });
}

int omx_attn_decode_fused_sharded_sync(const omx_array* out_full, const omx_array* q, const omx_array* k_new,
                                       const omx_array* v_new, omx_kv_cache cache, int rope_dims, bool traditional,
                                       omx_optional_float base, float rope_scale, const omx_array* freqs,
                                       float sm_scale, const omx_peer_group* peers, int head_offset, omx_stream s) {
  return guarded([&] {
    OMX_CHECK(peers != nullptr, "[attn_decode_fused_sharded] null peer group");
    decode_fused_impl(out_full, q, k_new, v_new, cache, rope_dims, traditional, base, rope_scale, freqs,
                      sm_scale, nullptr, nullptr, peers, head_offset, (cudaStream_t)s, nullptr, nullptr, 0.f,
                      /*peer_wait=*/true);
  });
}

int omx_attn_decode_fused_sharded_ll(const omx_array* out_full, const omx_array* q, const omx_array* k_new,
                                     const omx_array* v_new, omx_kv_cache cache, int rope_dims, bool traditional,
                                     omx_optional_float base, float rope_scale, const omx_array* freqs,
                                     float sm_scale, const omx_ll_group* group, int head_offset, omx_stream s) {
  return guarded([&] {
    OMX_CHECK(group != nullptr, "[attn_decode_fused_sharded_ll] null group");
    decode_fused_impl(out_full, q, k_new, v_new, cache, rope_dims, traditional, base, rope_scale, freqs,
                      sm_scale, nullptr, nullptr, nullptr, head_offset, (cudaStream_t)s, nullptr, nullptr, 0.f,
                      false, group);
  });
}

size_t omx_ll_staging_bytes(int world, int64_t B, int64_t Hq_local, int64_t D, int dtype) {
  if (world < 1 || B < 0 || Hq_local < 0 || D < 0) return 0;
  const size_t es = (dtype == OMX_FLOAT32) ? 4 : 2;
  return (size_t)2 * (size_t)world * ((size_t)B * Hq_local * D * es / 4) * 8;
}

int omx_attn_decode_seqshard(const omx_array* partial, const omx_array* q, const omx_array* k_new,
                             const omx_array* v_new, omx_kv_cache cache, int rope_dims, bool traditional,
                             omx_optional_float base, float rope_scale, int position, bool append, float sm_scale,
                             const omx_peer_group* peers, omx_stream s) {
  return guarded([&] {
    require_device();
    cudaStream_t stream = (cudaStream_t)s;
    auto* c = (KVCacheImpl*)cache.ctx;
    OMX_CHECK(c, "[attn_decode_seqshard] null cache handle");
    OMX_CHECK(q && partial, "[attn_decode_seqshard] null array");
    OMX_CHECK(q->ndim == 4 && q->shape[2] == 1, "[attn_decode_seqshard] q must be [B, Hq, 1, D]");
    const int64_t B = q->shape[0], Hq = q->shape[1], D = q->shape[3];
    const int world = peers ? peers->world : 1, rank = peers ? peers->rank : 0;
    OMX_CHECK(world >= 1 && world <= OMX_MAX_PEERS && rank >= 0 && rank < world, "[attn_decode_seqshard] bad peer group");
    OMX_CHECK(partial->dtype == OMX_FLOAT32 && partial->ndim == 4 && partial->shape[0] == world &&
                  partial->shape[1] == B && partial->shape[2] == Hq && partial->shape[3] == D + 2 &&
                  partial->strides[3] == 1 && partial->strides[2] == D + 2 && partial->strides[1] == Hq * (D + 2) &&
                  partial->strides[0] == B * Hq * (D + 2),
              "[attn_decode_seqshard] partial must be a contiguous float32 [world, B, Hq, D + 2] buffer");
    OMX_CHECK(position >= 0, "[attn_decode_seqshard] negative position");
    OMX_CHECK(rope_dims >= 0 && rope_dims % 2 == 0 && rope_dims <= D, "[rope] dims must be even and <= %lld", (long long)D);
    OMX_CHECK(rope_dims == 0 || base.has_value, "[attn_decode_seqshard] rope needs a base (no freqs here)");
    if (append) {
      OMX_CHECK(k_new && v_new && k_new->ndim == 4 && v_new->ndim == 4 && k_new->shape[2] == 1 && v_new->shape[2] == 1 &&
                    k_new->dtype == q->dtype && v_new->dtype == q->dtype,
                "[attn_decode_seqshard] the appending rank needs k_new / v_new [B, Hkv, 1, D] in q's dtype");
    }
    if (peers) {
      for (int r = 0; r < world; ++r)
        OMX_CHECK(peers->out[r] && peers->flags[r], "[attn_decode_seqshard] peer %d is not mapped", r);
      OMX_CHECK(peers->out[rank] == partial->data, "[attn_decode_seqshard] partial must be this rank's buffer of the peer group");
    }
    omx_array kview, vview;
    KVCacheTxn txn(c, stream);  // a failure below restores offset / capacity
    if (append) {
      kv_cache_update(c, k_new, v_new, &kview, &vview, /*skip_copy=*/true, stream);
    } else {
      kv_cache_state(c, &kview, &vview);
      kview.shape[2] = vview.shape[2] = kv_cache_offset(c);
    }
    OMX_CHECK(kview.shape[2] >= 1, "[attn_decode_seqshard] every rank must hold at least one key");
    // this rank's slot of the (local) partial buffer, described as the kernel's [B, Hq, 1, D] output
    const size_t slot_bytes = (size_t)B * Hq * (D + 2) * sizeof(float);
    omx_array slot = *q;
    slot.data = (char*)partial->data + (size_t)rank * slot_bytes;
    slot.strides[0] = Hq * (D + 2);
    slot.strides[1] = D + 2;
    slot.strides[2] = D + 2;
    slot.strides[3] = 1;
    SdpaArgs a = make_sdpa_args(&slot, q, &kview, &vview, sm_scale, "", nullptr, nullptr);
    const char* why = nullptr;
    const bool fast = decode_supported(a, &why) && (!append || (k_new->strides[3] == 1 && v_new->strides[3] == 1));
    OMX_CHECK(fast, "[attn_decode_seqshard] layout not supported by the decode kernels: %s",
              why ? why : "strided k_new/v_new");
    DecodeFused f;
    f.enabled = true;
    f.append = append;
    f.partial = true;
    f.k_new = k_new;
    f.v_new = v_new;
    f.rope_dims = rope_dims;
    f.traditional = traditional;
    f.position = position;  // GLOBAL position of the new token: rope row; the cache row is the local offset
    if (rope_dims > 0) f.table = get_rope_table(rope_dims, true, base.value, rope_scale, nullptr, position + 1, stream);
    omx_peer_group shifted;
    if (peers) {  // every rank's buffer receives this rank's slot
      shifted = *peers;
      for (int r = 0; r < world; ++r) shifted.out[r] = (char*)peers->out[r] + (size_t)rank * slot_bytes;
      f.peers = &shifted;
    }
    decode_attention(a, f, stream);
    txn.commit();
  });
}

int omx_seqshard_merge(const omx_array* out, const omx_array* partial, const omx_peer_group* peers, uint32_t expected,
                       omx_stream s) {
  return guarded([&] {
    require_device();
    OMX_CHECK(out && partial && out->ndim == 4 && partial->ndim == 4 && out->shape[2] == 1, "[seqshard_merge] bad arrays");
    const int64_t world = partial->shape[0], B = out->shape[0], Hq = out->shape[1], D = out->shape[3];
    OMX_CHECK(partial->dtype == OMX_FLOAT32 && partial->shape[1] == B && partial->shape[2] == Hq &&
                  partial->shape[3] == D + 2 && partial->strides[3] == 1 && partial->strides[2] == D + 2 &&
                  partial->strides[1] == Hq * (D + 2) && partial->strides[0] == B * Hq * (D + 2),
              "[seqshard_merge] partial must be a contiguous float32 [world, B, Hq, D + 2] buffer");
    OMX_CHECK(is_float_dtype(out->dtype), "[seqshard_merge] out must be floating point");
    OMX_CHECK(world >= 1 && world <= OMX_MAX_PEERS && (!peers || (peers->world == world && peers->rank >= 0 &&
                  peers->rank < world && peers->flags[peers->rank])), "[seqshard_merge] bad peer group");
    seqshard_merge(out, (const float*)partial->data, (int)world, (int)B, (int)Hq, (int)D,
                   peers ? peers->flags[peers->rank] : nullptr, expected, peers ? peers->rank : 0, (cudaStream_t)s);
  });
}

int omx_peer_wait(const omx_peer_group* peers, uint32_t expected, omx_stream s) {
  return guarded([&] {
    require_device();
    OMX_CHECK(peers && peers->world >= 1 && peers->world <= OMX_MAX_PEERS && peers->rank >= 0 &&
                  peers->rank < peers->world && peers->flags[peers->rank],
              "[peer_wait] bad peer group");
    peer_wait(peers->flags[peers->rank], peers->world, expected, peers->rank, (cudaStream_t)s);
  });
}

int omx_dit_rope(const omx_array* out, const omx_array* x, const omx_array* cos, const omx_array* sin,
                 omx_stream s) {
  return guarded([&] {
    require_device();
    dit_rope_forward(out, x, cos, sin, (cudaStream_t)s);
  });
}

// softmax(scale q k^T [+ mask]) v for the DiT callers.  The reference chains promote to f32 after the first
// matmul (f32 mask, f32 output for 16-bit inputs): the tcgen05 kernel takes both (f32 epilogue, f32 mask rows
// in its mixed tiles); whatever it refuses goes to the generic kernel.
static void dit_attention_impl(const omx_array* out, const omx_array* q, const omx_array* k, const omx_array* v,
                               float scale, const omx_array* add_mask, cudaStream_t stream) {
  const bool has_mask = add_mask && add_mask->data;
  if (has_mask)
    OMX_CHECK(add_mask->dtype == OMX_FLOAT32 || add_mask->dtype == q->dtype,
              "[dit_joint_attention] add_mask must be float32 or the input dtype");
  SdpaArgs a{};
  if (has_mask && add_mask->dtype == OMX_FLOAT32 && q->dtype != OMX_FLOAT32) {
    a = make_sdpa_args(out, q, k, v, scale, "", nullptr, nullptr);
    a.mask_mode = MASK_ADD;
    a.mask = add_mask;
    const int64_t full[4] = {a.B, a.Hq, a.Lq, a.Lk};
    const int lead = 4 - add_mask->ndim;
    OMX_CHECK(add_mask->ndim <= 4, "[dit_joint_attention] mask rank > 4");
    for (int i = 0; i < 4; ++i) {
      if (i < lead) { a.mask_strides[i] = 0; continue; }
      const int64_t n = add_mask->shape[i - lead];
      OMX_CHECK(n == full[i] || n == 1, "[dit_joint_attention] mask not broadcastable");
      a.mask_strides[i] = (n == 1 && full[i] != 1) ? 0 : add_mask->strides[i - lead];
    }
    if ((int64_t)a.B * a.Hq * a.Lq * a.Dv == 0) return;
    const char* why = nullptr;
    const bool forced_generic = t_forced_kernel == "sdpa_generic";
    if (!forced_generic && fmha_sm100_supported(a, &why) && (a.Lq >= 2 || t_forced_kernel == "fmha_tcgen05")) {
      fmha_sm100(a, stream);
      return;
    }
    if (!forced_generic && t_forced_kernel != "fmha_tcgen05" && sdpa_f32_tiled_supported(a, &why) &&
        (a.Lq >= 16 || t_forced_kernel == "sdpa_f32_tiled")) {
      sdpa_f32_tiled(a, stream);  // the float32 DiT chains (the FLUX example runs in float32)
      return;
    }
    OMX_CHECK(t_forced_kernel.empty() || forced_generic, "forced kernel '%s' does not support this call: %s",
              t_forced_kernel.c_str(), why ? why : "?");
    sdpa_generic(a, stream);
    return;
  }
  a = make_sdpa_args(out, q, k, v, scale, "", has_mask ? add_mask : nullptr, nullptr);
  dispatch_sdpa(a, stream);
}

int omx_dit_joint_attention(const omx_array* out, const omx_array* q, const omx_array* k, const omx_array* v,
                            float scale, const omx_array* add_mask, omx_stream s) {
  return guarded([&] {
    require_device();
    dit_attention_impl(out, q, k, v, scale, add_mask, (cudaStream_t)s);
  });
}

// [B,S,H,D] storage -> the [B,H,S,D] view the kernels take
static omx_array bshd_as_bhsd(const omx_array& a) {
  omx_array t = a;
  t.shape[1] = a.shape[2]; t.shape[2] = a.shape[1];
  t.strides[1] = a.strides[2]; t.strides[2] = a.strides[1];
  return t;
}
// rows [t0, t0 + n) along axis 1 of a [B,S,...] array
static omx_array rows_axis1(const omx_array& a, int64_t t0, int64_t n) {
  omx_array t = a;
  t.shape[1] = n;
  t.data = (char*)a.data + (size_t)t0 * a.strides[1] * dtype_size(a.dtype);
  return t;
}

int omx_dit_attn_fused(const omx_array* out, int n_streams, const omx_array* const* q, const omx_array* const* k,
                       const omx_array* const* v, const omx_array* const* q_norm_weight,
                       const omx_array* const* k_norm_weight, float norm_eps, const omx_array* cos,
                       const omx_array* sin, float scale, const omx_array* add_mask, omx_stream s) {
  return guarded([&] {
    require_device();
    cudaStream_t stream = (cudaStream_t)s;
    OMX_CHECK(n_streams == 1 || n_streams == 2, "[dit_attn_fused] n_streams must be 1 or 2, got %d", n_streams);
    OMX_CHECK(out && q && k && v, "[dit_attn_fused] null array list");
    int64_t S = 0;
    for (int i = 0; i < n_streams; ++i) {
      OMX_CHECK(q[i] && k[i] && v[i], "[dit_attn_fused] null array in stream %d", i);
      OMX_CHECK(q[i]->ndim == 4 && k[i]->ndim == 4 && v[i]->ndim == 4, "[dit_attn_fused] q, k, v must be [B,S,H,D]");
      OMX_CHECK(is_float_dtype(q[i]->dtype) && k[i]->dtype == q[0]->dtype && v[i]->dtype == q[0]->dtype &&
                    q[i]->dtype == q[0]->dtype,
                "[dit_attn_fused] q, k, v of all streams must share one floating dtype");
      OMX_CHECK(q[i]->shape[0] == q[0]->shape[0] && k[i]->shape[0] == q[0]->shape[0] &&
                    v[i]->shape[0] == q[0]->shape[0],
                "[dit_attn_fused] mismatching batch dimension");
      OMX_CHECK(q[i]->shape[1] == k[i]->shape[1] && q[i]->shape[1] == v[i]->shape[1],
                "[dit_attn_fused] q, k, v of stream %d must have the same number of tokens", i);
      OMX_CHECK(q[i]->shape[2] == q[0]->shape[2] && k[i]->shape[2] == k[0]->shape[2] &&
                    v[i]->shape[2] == k[0]->shape[2] && q[i]->shape[3] == q[0]->shape[3] &&
                    k[i]->shape[3] == q[0]->shape[3] && v[i]->shape[3] == v[0]->shape[3],
                "[dit_attn_fused] head counts / head dims differ between streams");
      S += q[i]->shape[1];
    }
    const int dt = q[0]->dtype;
    const int64_t B = q[0]->shape[0], H = q[0]->shape[2], Hkv = k[0]->shape[2], D = q[0]->shape[3],
                  Dv = v[0]->shape[3];
    OMX_CHECK(out->ndim == 4 && out->shape[0] == B && out->shape[1] == S && out->shape[2] == H &&
                  out->shape[3] == Dv,
              "[dit_attn_fused] out must be [B, S_total, H, Dv]");
    const bool has_rope = cos && cos->data && sin && sin->data;
    int64_t tcs[3] = {0, 0, 0}, tss[3] = {0, 0, 0};
    if (has_rope) {
      OMX_CHECK(D % 2 == 0, "[dit_attn_fused] head_dim must be even");
      auto tbl = [&](const omx_array* a, int64_t st[3], const char* nm) {
        OMX_CHECK(a->dtype == dt, "[dit_attn_fused] %s must have the q dtype", nm);
        if (a->ndim == 3) {
          OMX_CHECK(a->shape[0] == B && a->shape[1] == S && a->shape[2] == D / 2,
                    "[dit_attn_fused] %s must be [B, S_total, D/2]", nm);
          st[0] = a->strides[0]; st[1] = a->strides[1]; st[2] = a->strides[2];
        } else {
          OMX_CHECK(a->ndim == 4 && a->shape[0] == B && a->shape[1] == S && a->shape[2] == 1 &&
                        a->shape[3] == D / 2,
                    "[dit_attn_fused] %s must be [B, S_total, D/2] or [B, S_total, 1, D/2]", nm);
          st[0] = a->strides[0]; st[1] = a->strides[1]; st[2] = a->strides[3];
        }
      };
      tbl(cos, tcs, "cos");
      tbl(sin, tss, "sin");
    }
    auto wt = [&](const omx_array* const* list, int i) -> const omx_array* {
      return (list && list[i] && list[i]->data) ? list[i] : nullptr;
    };
    bool any_norm = false;
    for (int i = 0; i < n_streams; ++i) any_norm = any_norm || wt(q_norm_weight, i) || wt(k_norm_weight, i);
    if (B * S * H == 0) return;
    // joint [txt; img] buffers (K/V order = the order of the streams, klein_model.rs:461-462)
    const size_t es = dtype_size(dt);
    auto al = [](size_t n) { return (n + 255) & ~(size_t)255; };
    const bool in_place_qk = n_streams == 1 && !any_norm && !has_rope;
    const bool in_place_v = n_streams == 1;
    const size_t qb = in_place_qk ? 0 : al((size_t)B * S * H * D * es);
    const size_t kb = in_place_qk ? 0 : al((size_t)B * S * Hkv * D * es);
    const size_t vb = in_place_v ? 0 : al((size_t)B * S * Hkv * Dv * es);
    char* ws = (qb + kb + vb) ? (char*)get_outer_workspace(qb + kb + vb, stream) : nullptr;
    auto dense = [&](char* mem, int64_t heads, int64_t d) {
      omx_array t{};
      t.data = mem; t.dtype = dt; t.ndim = 4;
      t.shape[0] = B; t.shape[1] = S; t.shape[2] = heads; t.shape[3] = d;
      t.strides[3] = 1; t.strides[2] = d; t.strides[1] = heads * d; t.strides[0] = S * heads * d;
      return t;
    };
    omx_array Qj = in_place_qk ? *q[0] : dense(ws, H, D);
    omx_array Kj = in_place_qk ? *k[0] : dense(ws + qb, Hkv, D);
    omx_array Vj = in_place_v ? *v[0] : dense(ws + qb + kb, Hkv, Dv);
    if (!in_place_qk || !in_place_v) {
      PrologueCall pc;
      pc.mode = 2;
      pc.dims = (int)D;
      pc.eps = norm_eps;
      pc.tcos = has_rope ? cos : nullptr;
      pc.tsin = has_rope ? sin : nullptr;
      for (int j = 0; j < 3; ++j) { pc.tcs[j] = tcs[j]; pc.tss[j] = tss[j]; }
      omx_array xin[6];
      int n = 0;
      int64_t t0 = 0;
      for (int i = 0; i < n_streams; ++i) {
        const int64_t Si = q[i]->shape[1];
        if (!in_place_qk) {
          xin[n] = bshd_as_bhsd(*q[i]);
          pc.seg[n].x = &xin[n]; pc.seg[n].out = bshd_as_bhsd(rows_axis1(Qj, t0, Si));
          pc.seg[n].w = wt(q_norm_weight, i); pc.seg[n].rope = has_rope; pc.seg[n].tok0 = (int)t0; ++n;
          xin[n] = bshd_as_bhsd(*k[i]);
          pc.seg[n].x = &xin[n]; pc.seg[n].out = bshd_as_bhsd(rows_axis1(Kj, t0, Si));
          pc.seg[n].w = wt(k_norm_weight, i); pc.seg[n].rope = has_rope; pc.seg[n].tok0 = (int)t0; ++n;
        }
        if (!in_place_v && Dv == D) {
          xin[n] = bshd_as_bhsd(*v[i]);
          pc.seg[n].x = &xin[n]; pc.seg[n].out = bshd_as_bhsd(rows_axis1(Vj, t0, Si));
          pc.seg[n].w = nullptr; pc.seg[n].rope = false; ++n;
        }
        t0 += Si;
      }
      pc.nseg = n;
      note_launch("qkv_prologue");
      const bool fused = n > 0 && qkv_prologue(pc, stream);
      t0 = 0;
      for (int i = 0; i < n_streams; ++i) {  // whatever the one-launch kernel did not cover
        const int64_t Si = q[i]->shape[1];
        if (!fused && !in_place_qk) {
          const omx_array* src[2] = {q[i], k[i]};
          omx_array dst[2] = {rows_axis1(Qj, t0, Si), rows_axis1(Kj, t0, Si)};
          const omx_array* w[2] = {wt(q_norm_weight, i), wt(k_norm_weight, i)};
          for (int j = 0; j < 2; ++j) {
            const omx_array* cur = src[j];
            if (w[j]) {
              rms_norm_forward(&dst[j], cur, w[j], norm_eps, stream);
              cur = &dst[j];
            }
            if (has_rope) {
              omx_array c = rows_axis1(*cos, t0, Si), sn = rows_axis1(*sin, t0, Si);
              dit_rope_forward(&dst[j], cur, &c, &sn, stream);
            } else if (!w[j]) {
              omx_array d4 = bshd_as_bhsd(dst[j]), s4 = bshd_as_bhsd(*cur);
              copy4d(&d4, &s4, stream);
            }
          }
        }
        if (!in_place_v && (!fused || Dv != D)) {
          omx_array d4 = bshd_as_bhsd(rows_axis1(Vj, t0, Si)), s4 = bshd_as_bhsd(*v[i]);
          copy4d(&d4, &s4, stream);
        }
        t0 += Si;
      }
    }
    omx_array qv = bshd_as_bhsd(Qj), kv = bshd_as_bhsd(Kj), vv = bshd_as_bhsd(Vj), ov = bshd_as_bhsd(*out);
    dit_attention_impl(&ov, &qv, &kv, &vv, scale, add_mask, stream);
  });
}

}  // extern "C"
