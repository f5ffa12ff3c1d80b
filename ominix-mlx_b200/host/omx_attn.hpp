// omx_attn.hpp -- C++ host layer over the C ABI (include/omx_attn.h), header-only.
//
// The reference's host code for this path is Rust (mlx-rs / mlx-rs-core); the image has no Rust
// toolchain, so the compiled-language mirror of that layer is C++: same names, argument order and
// error behaviour as the Rust API it stands in for --
//   omx::fast::rope / rms_norm / scaled_dot_product_attention   <- mlx-rs/src/fast.rs:15-46,110-180
//   omx::fast::ScaledDotProductAttentionMask {Array, Arrays, Causal}   <- fast.rs:53-108
//   omx::nn::Rope / RmsNorm                                     <- mlx-rs/src/nn/positional_encoding.rs:17-137,
//                                                                  nn/normalization.rs:209-270
//   omx::KeyValueCache / KVCache / ConcatKeyValueCache           <- mlx-rs-core/src/cache.rs:7-195
//   omx::utils::{initialize_rope, SdpaMask, scaled_dot_product_attention}   <- mlx-rs-core/src/utils.rs:52-209
//   omx::utils::{attention_decode_fused, attention_prefill_fused}  <- Attention::forward, qwen3-mlx/src/model.rs:161-215
//   omx::PagedKVCache, omx::utils::attention_decode_fused_paged    <- the same contract over a page pool (ragged batches)
// Errors: every failing call throws omx::Exception{what} carrying the library message (the role
// mlx_rs::error::Exception plays, mlx-rs/src/error.rs:236-288).
// `Array` is a strided view of CUDA device memory (what mlx_rs::Array is on this path): owning when
// made by Array::empty / from_host, borrowed when it wraps caller memory or a cache-owned buffer.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/omx_attn.h"

namespace omx {

struct Exception : public std::runtime_error {
  using std::runtime_error::runtime_error;
  const char* what_() const { return what(); }
};

inline void check(int status) {
  if (status != 0) throw Exception(omx_last_error());
}
inline void cuda_check(cudaError_t e, const char* what) {
  if (e != cudaSuccess) throw Exception(std::string(what) + ": " + cudaGetErrorString(e));
}

enum class Dtype : int32_t { Bool = OMX_BOOL, Int32 = OMX_INT32, Float16 = OMX_FLOAT16, Float32 = OMX_FLOAT32,
                             Bfloat16 = OMX_BFLOAT16 };
inline size_t dtype_size(Dtype d) {
  switch (d) {
    case Dtype::Bool: return 1;
    case Dtype::Float16: case Dtype::Bfloat16: return 2;
    default: return 4;
  }
}

/// A cudaStream_t; the default-constructed value is the legacy default stream (the analogue of
/// Stream::task_local_or_default() that #[default_device] injects, mlx-internal-macros/src/lib.rs:83-94).
struct Stream {
  cudaStream_t s = nullptr;
  omx_stream raw() const { return (omx_stream)s; }
};

class Array {
 public:
  Array() { std::memset(&d_, 0, sizeof(d_)); }

  /// Fresh contiguous device array (uninitialised).
  static Array empty(const std::vector<int64_t>& shape, Dtype dtype) {
    Array a;
    a.d_.dtype = (int32_t)dtype;
    a.d_.ndim = (int32_t)shape.size();
    int64_t n = 1;
    for (int i = (int)shape.size() - 1; i >= 0; --i) {
      a.d_.shape[i] = shape[i];
      a.d_.strides[i] = n;
      n *= shape[i];
    }
    void* p = nullptr;
    cuda_check(cudaMalloc(&p, (size_t)std::max<int64_t>(n, 1) * dtype_size(dtype)), "cudaMalloc");
    a.d_.data = p;
    a.own_ = std::shared_ptr<void>(p, [](void* q) { cudaFree(q); });
    return a;
  }
  /// Upload `count * dtype_size` bytes of host data into a fresh array of `shape`.
  static Array from_host(const void* host, const std::vector<int64_t>& shape, Dtype dtype, Stream s = {}) {
    Array a = empty(shape, dtype);
    cuda_check(cudaMemcpyAsync(a.d_.data, host, (size_t)a.size() * dtype_size(dtype), cudaMemcpyHostToDevice, s.s),
               "cudaMemcpyAsync H2D");
    return a;
  }
  /// Borrow memory owned elsewhere (a projection output of the host framework); strides in ELEMENTS.
  static Array borrow(void* data, Dtype dtype, const std::vector<int64_t>& shape, const std::vector<int64_t>& strides) {
    if (shape.size() != strides.size() || shape.size() > OMX_MAX_NDIM) throw Exception("bad shape/strides");
    Array a;
    a.d_.data = data;
    a.d_.dtype = (int32_t)dtype;
    a.d_.ndim = (int32_t)shape.size();
    for (size_t i = 0; i < shape.size(); ++i) {
      a.d_.shape[i] = shape[i];
      a.d_.strides[i] = strides[i];
    }
    return a;
  }
  /// View of library-owned memory (cache buffers); `keep` pins the owner.
  static Array from_desc(const omx_array& d, std::shared_ptr<void> keep) {
    Array a;
    a.d_ = d;
    a.own_ = std::move(keep);
    return a;
  }

  std::vector<int64_t> shape() const { return std::vector<int64_t>(d_.shape, d_.shape + d_.ndim); }
  std::vector<int64_t> strides() const { return std::vector<int64_t>(d_.strides, d_.strides + d_.ndim); }
  int ndim() const { return d_.ndim; }
  Dtype dtype() const { return (Dtype)d_.dtype; }
  int64_t size() const {
    int64_t n = 1;
    for (int i = 0; i < d_.ndim; ++i) n *= d_.shape[i];
    return n;
  }
  void* data() const { return d_.data; }
  const omx_array* desc() const { return &d_; }
  bool is_contiguous() const {
    int64_t n = 1;
    for (int i = d_.ndim - 1; i >= 0; --i) {
      if (d_.shape[i] != 1 && d_.strides[i] != n) return false;
      n *= d_.shape[i];
    }
    return true;
  }

  /// transpose_axes(&[0, 2, 1, 3]): permutes shape and strides, no copy (how the crates turn their
  /// [B, L, H, D] projections into [B, H, L, D], qwen3-mlx/src/model.rs:172-184).
  Array transpose_axes(const std::vector<int>& axes) const {
    if ((int)axes.size() != d_.ndim) throw Exception("transpose_axes: wrong number of axes");
    Array a = *this;
    for (int i = 0; i < d_.ndim; ++i) {
      a.d_.shape[i] = d_.shape[axes[i]];
      a.d_.strides[i] = d_.strides[axes[i]];
    }
    return a;
  }
  /// Download a CONTIGUOUS array (synchronises the stream).
  void to_host(void* host, Stream s = {}) const {
    if (!is_contiguous()) throw Exception("to_host needs a contiguous array");
    cuda_check(cudaMemcpyAsync(host, d_.data, (size_t)size() * dtype_size(dtype()), cudaMemcpyDeviceToHost, s.s),
               "cudaMemcpyAsync D2H");
    cuda_check(cudaStreamSynchronize(s.s), "cudaStreamSynchronize");
  }

 private:
  omx_array d_;
  std::shared_ptr<void> own_;
};

// ------------------------------------------------------------------------------- fast
namespace fast {

inline omx_optional_float opt(std::optional<float> v) {
  omx_optional_float o;
  o.has_value = v.has_value();
  o.value = v.value_or(0.f);
  return o;
}

/// fast::rope (mlx-rs/src/fast.rs:15-46): exactly one of `base` / `freqs`.
inline Array rope(const Array& array, int dimensions, bool traditional, std::optional<float> base, float scale,
                  int offset, const Array* freqs = nullptr, Stream s = {}) {
  Array out = Array::empty(array.shape(), array.dtype());
  check(omx_fast_rope(out.desc(), array.desc(), dimensions, traditional, opt(base), scale, offset,
                      freqs ? freqs->desc() : nullptr, s.raw()));
  return out;
}

/// fast::rms_norm (mlx-rs/src/fast.rs:163-180).
inline Array rms_norm(const Array& x, const Array& weight, float eps, Stream s = {}) {
  Array out = Array::empty(x.shape(), x.dtype());
  check(omx_fast_rms_norm(out.desc(), x.desc(), weight.desc(), eps, s.raw()));
  return out;
}

/// fast.rs:53-62.  Arrays(...) only uses its first entry (fast.rs:95-102).
struct ScaledDotProductAttentionMask {
  enum Kind { None, ArrayMask, Causal } kind = None;
  const omx::Array* array = nullptr;
  static ScaledDotProductAttentionMask none() { return {}; }
  static ScaledDotProductAttentionMask causal() { return {Causal, nullptr}; }
  static ScaledDotProductAttentionMask from(const omx::Array& m) { return {ArrayMask, &m}; }
  static ScaledDotProductAttentionMask from(const std::vector<omx::Array>& ms) {
    return ms.empty() ? ScaledDotProductAttentionMask{} : ScaledDotProductAttentionMask{ArrayMask, &ms[0]};
  }
  const char* mode() const { return kind == Causal ? "causal" : ""; }
  const omx_array* ptr() const { return kind == ArrayMask && array ? array->desc() : nullptr; }
};

/// fast::scaled_dot_product_attention (mlx-rs/src/fast.rs:110-151): O = softmax(scale Q K^T + mask) V,
/// GQA without pre-tiling, f32 softmax; output [B, Hq, Lq, Dv] in the input dtype.
inline Array scaled_dot_product_attention(const Array& queries, const Array& keys, const Array& values, float scale,
                                          ScaledDotProductAttentionMask mask = {}, Stream s = {},
                                          Array* out_into = nullptr) {
  Array out;
  if (out_into) {
    out = *out_into;
  } else {
    std::vector<int64_t> sh = queries.shape();
    if (queries.ndim() == 4 && values.ndim() == 4) sh[3] = values.shape()[3];
    out = Array::empty(sh, queries.dtype());
  }
  check(omx_fast_scaled_dot_product_attention(out.desc(), queries.desc(), keys.desc(), values.desc(), scale,
                                              mask.mode(), mask.ptr(), nullptr, s.raw()));
  return out;
}

}  // namespace fast

// ------------------------------------------------------------------------------- nn
namespace nn {

/// nn::Rope = RotaryPositionalEncoding {dimensions, traditional = false, base = 10000, scale = 1}
/// (positional_encoding.rs:57-66); forward((x, offset)) = fast::rope(...) (:120-134).
struct Rope {
  int dimensions;
  bool traditional = false;
  float base = 10000.f;
  float scale = 1.f;
  Array forward(const Array& x, int offset = 0, Stream s = {}) const {
    return fast::rope(x, dimensions, traditional, base, scale, offset, nullptr, s);
  }
};

/// nn::RmsNorm {weight, eps = 1e-5} (normalization.rs:209-270).
struct RmsNorm {
  Array weight;
  float eps = 1e-5f;
  Array forward(const Array& x, Stream s = {}) const { return fast::rms_norm(x, weight, eps, s); }
};

}  // namespace nn

// ------------------------------------------------------------------------------- caches
/// trait KeyValueCache (cache.rs:7-20).
class KeyValueCache {
 public:
  virtual ~KeyValueCache() = default;
  virtual int offset() const = 0;
  virtual std::optional<int> max_size() const = 0;
  virtual std::pair<Array, Array> update_and_fetch(const Array& keys, const Array& values) = 0;
  virtual void reset() {}  // trait default: does nothing (cache.rs:17-19)
};

/// KVCache {keys, values, offset, step = 256} (cache.rs:92-195).
class KVCache : public KeyValueCache {
 public:
  explicit KVCache(int step = 256, Stream s = {}) : stream_(s) {
    omx_kv_cache h{nullptr};
    check(omx_kv_cache_new(&h, step));
    h_ = std::shared_ptr<void>(h.ctx, [](void* c) { omx_kv_cache_free(omx_kv_cache{c}); });
  }
  static KVCache with_step(int step) { return KVCache(step); }
  int offset() const override {
    int n = 0;
    check(omx_kv_cache_offset(raw(), &n));
    return n;
  }
  std::optional<int> max_size() const override { return std::nullopt; }
  void reset() override { check(omx_kv_cache_reset(raw())); }
  std::pair<Array, Array> update_and_fetch(const Array& keys, const Array& values) override {
    omx_array ko, vo;
    check(omx_kv_cache_update_and_fetch(raw(), keys.desc(), values.desc(), &ko, &vo, stream_.raw()));
    return {Array::from_desc(ko, h_), Array::from_desc(vo, h_)};
  }
  /// The whole backing buffers [B, Hkv, cap, D] (self.keys / self.values in the reference).
  std::pair<Array, Array> state() const {
    omx_array ko, vo;
    check(omx_kv_cache_state(raw(), &ko, &vo));
    return {Array::from_desc(ko, h_), Array::from_desc(vo, h_)};
  }
  /// CUDA-graph decode loop (include/omx_attn.h): pin buffers + split-K scratch for positions [0, max_rows).
  void prepare_graph(int max_rows, int n_q_heads) { check(omx_kv_cache_prepare_graph(raw(), max_rows, n_q_heads, stream_.raw())); }
  /// Host bookkeeping for n rows appended by dynamic-position launches / graph replays.
  void advance(int n = 1) { check(omx_kv_cache_advance(raw(), n, stream_.raw())); }
  omx_kv_cache raw() const { return omx_kv_cache{h_.get()}; }
  std::shared_ptr<void> keepalive() const { return h_; }

 private:
  std::shared_ptr<void> h_;
  Stream stream_;
};

/// ConcatKeyValueCache (cache.rs:45-85).
class ConcatKeyValueCache : public KeyValueCache {
 public:
  explicit ConcatKeyValueCache(Stream s = {}) : stream_(s) {
    omx_kv_cache h{nullptr};
    check(omx_concat_kv_cache_new(&h));
    h_ = std::shared_ptr<void>(h.ctx, [](void* c) { omx_concat_kv_cache_free(omx_kv_cache{c}); });
  }
  int offset() const override {
    int n = 0;
    check(omx_concat_kv_cache_offset(omx_kv_cache{h_.get()}, &n));
    return n;
  }
  std::optional<int> max_size() const override { return std::nullopt; }
  std::pair<Array, Array> update_and_fetch(const Array& keys, const Array& values) override {
    omx_array ko, vo;
    check(omx_concat_kv_cache_update_and_fetch(omx_kv_cache{h_.get()}, keys.desc(), values.desc(), &ko, &vo,
                                               stream_.raw()));
    return {Array::from_desc(ko, h_), Array::from_desc(vo, h_)};
  }

 private:
  std::shared_ptr<void> h_;
  Stream stream_;
};

/// The KeyValueCache contract over a page pool (include/omx_attn.h "paged KV cache"): pool [n_pages][Hkv][64][D]
/// allocated once, block table, per-sequence lengths; growth = a page id from the free list.  update_and_fetch
/// appends to EVERY sequence and returns MATERIALISED [B,Hkv,offset,D] arrays (valid until the next fetch).
class PagedKVCache : public KeyValueCache {
 public:
  PagedKVCache(int batch, int n_kv_heads, int head_dim, Dtype dtype, int64_t n_pages, int max_pages_per_seq,
               int head_dim_v = 0, Stream s = {})
      : batch_(batch), stream_(s) {
    omx_paged_kv_cache h{nullptr};
    check(omx_paged_kv_cache_new(&h, batch, n_kv_heads, head_dim, head_dim_v ? head_dim_v : head_dim, (int)dtype,
                                 n_pages, max_pages_per_seq));
    h_ = std::shared_ptr<void>(h.ctx, [](void* c) { omx_paged_kv_cache_free(omx_paged_kv_cache{c}); });
  }
  int offset() const override {  // the longest sequence
    int n = 0;
    check(omx_paged_kv_cache_offset(raw(), &n));
    return n;
  }
  std::optional<int> max_size() const override { return std::nullopt; }
  void reset() override { check(omx_paged_kv_cache_reset(raw(), -1, stream_.raw())); }
  void reset(int slot) { check(omx_paged_kv_cache_reset(raw(), slot, stream_.raw())); }
  void release(int slot) { check(omx_paged_kv_cache_release(raw(), slot, stream_.raw())); }
  void reserve(int rows_ahead) { check(omx_paged_kv_cache_reserve(raw(), rows_ahead, stream_.raw())); }
  void trim(int n) { check(omx_paged_kv_cache_trim(raw(), n, stream_.raw())); }
  void sync_lengths() { check(omx_paged_kv_cache_sync_lengths(raw(), stream_.raw())); }
  std::vector<int32_t> lengths() const {
    std::vector<int32_t> l(batch_);
    check(omx_paged_kv_cache_lengths(raw(), l.data()));
    return l;
  }
  int64_t free_pages() const {
    int64_t n = 0;
    check(omx_paged_kv_cache_free_pages(raw(), &n));
    return n;
  }
  std::pair<Array, Array> update_and_fetch(const Array& keys, const Array& values) override {
    omx_array ko, vo;
    check(omx_paged_kv_cache_update_and_fetch(raw(), keys.desc(), values.desc(), &ko, &vo, stream_.raw()));
    return {Array::from_desc(ko, h_), Array::from_desc(vo, h_)};
  }
  /// append without materialising the fetched views
  void update(const Array& keys, const Array& values) {
    check(omx_paged_kv_cache_update_and_fetch(raw(), keys.desc(), values.desc(), nullptr, nullptr, stream_.raw()));
  }
  void append_slot(int slot, const Array& keys, const Array& values) {
    check(omx_paged_kv_cache_append_slot(raw(), slot, keys.desc(), values.desc(), stream_.raw()));
  }
  std::pair<Array, Array> fetch() {
    omx_array ko, vo;
    check(omx_paged_kv_cache_fetch(raw(), &ko, &vo, stream_.raw()));
    return {Array::from_desc(ko, h_), Array::from_desc(vo, h_)};
  }
  omx_paged_kv_cache raw() const { return omx_paged_kv_cache{h_.get()}; }

 private:
  std::shared_ptr<void> h_;
  int batch_;
  Stream stream_;
};

// ------------------------------------------------------------------------------- utils
namespace utils {

/// initialize_rope (utils.rs:52-97): rope_scaling type / rope_type in {default, linear}; linear => 1/factor.
inline nn::Rope initialize_rope(int dims, float base, bool traditional,
                                const std::map<std::string, std::string>* scaling_config = nullptr,
                                int /*max_position_embeddings*/ = 0) {
  std::string type = "default";
  if (scaling_config) {
    auto it = scaling_config->find("type");
    if (it == scaling_config->end()) it = scaling_config->find("rope_type");
    if (it != scaling_config->end()) type = it->second;
  }
  if (type != "default" && type != "linear") throw Exception("Unsupported RoPE type " + type);
  float scale = 1.f;
  if (type == "linear" && scaling_config) {
    auto it = scaling_config->find("factor");
    if (it == scaling_config->end()) throw Exception("key \"factor\" is not found in scaling config");
    try {
      scale = 1.f / std::stof(it->second);
    } catch (...) {
      throw Exception("key \"factor\" is not a valid float");
    }
  }
  return nn::Rope{dims, traditional, base, scale};
}

/// SdpaMask (utils.rs:105-116).
struct SdpaMask {
  enum Kind { None, Causal, ArrayMask } kind = None;
  const Array* array = nullptr;
  static SdpaMask none() { return {}; }
  static SdpaMask causal() { return {Causal, nullptr}; }
  static SdpaMask from(const Array& m) { return {ArrayMask, &m}; }
  fast::ScaledDotProductAttentionMask lower() const {
    if (kind == Causal) return fast::ScaledDotProductAttentionMask::causal();
    if (kind == ArrayMask) return fast::ScaledDotProductAttentionMask::from(*array);
    return {};
  }
};

/// scaled_dot_product_attention::<C>(q, k, v, Option<C>, scale, Option<SdpaMask>) (utils.rs:191-209);
/// the cache argument is unused there as well.
inline Array scaled_dot_product_attention(const Array& queries, const Array& keys, const Array& values,
                                          const KeyValueCache* /*cache*/, float scale, SdpaMask mask = {},
                                          Stream s = {}) {
  return fast::scaled_dot_product_attention(queries, keys, values, scale, mask.lower(), s);
}

/// The decode step of Attention::forward (L == 1) in ONE launch:
/// [q_norm, k_norm] -> rope(q, off), rope(k, off) -> cache.update_and_fetch -> sdpa(mask none).
inline Array attention_decode_fused(const Array& queries, const Array& keys, const Array& values, KVCache& cache,
                                    const nn::Rope* rope, float scale, const nn::RmsNorm* q_norm = nullptr,
                                    const nn::RmsNorm* k_norm = nullptr, Stream s = {}) {
  const auto qs = queries.shape();
  if (qs.size() != 4 || qs[2] != 1) throw Exception("attention_decode_fused: queries must be [B, H, 1, D]");
  Array out = Array::empty({qs[0], qs[1], 1, values.shape()[3]}, queries.dtype());
  const float eps = q_norm ? q_norm->eps : (k_norm ? k_norm->eps : 0.f);
  check(omx_attn_decode_fused_norm(out.desc(), queries.desc(), keys.desc(), values.desc(), cache.raw(),
                                   q_norm ? q_norm->weight.desc() : nullptr, k_norm ? k_norm->weight.desc() : nullptr,
                                   eps, rope ? rope->dimensions : 0, rope ? rope->traditional : false,
                                   fast::opt(rope ? std::optional<float>(rope->base) : std::nullopt),
                                   rope ? rope->scale : 1.f, nullptr, scale, nullptr, nullptr, s.raw()));
  return out;
}

/// The same step over a PagedKVCache: per sequence the position is its own length (device memory), ONE launch.
inline Array attention_decode_fused_paged(const Array& queries, const Array& keys, const Array& values,
                                          PagedKVCache& cache, const nn::Rope* rope, float scale,
                                          const nn::RmsNorm* q_norm = nullptr, const nn::RmsNorm* k_norm = nullptr,
                                          Stream s = {}) {
  const auto qs = queries.shape();
  if (qs.size() != 4 || qs[2] != 1) throw Exception("attention_decode_fused_paged: queries must be [B, H, 1, D]");
  Array out = Array::empty({qs[0], qs[1], 1, values.shape()[3]}, queries.dtype());
  const float eps = q_norm ? q_norm->eps : (k_norm ? k_norm->eps : 0.f);
  check(omx_attn_decode_fused_paged(out.desc(), queries.desc(), keys.desc(), values.desc(), cache.raw(),
                                    q_norm ? q_norm->weight.desc() : nullptr, k_norm ? k_norm->weight.desc() : nullptr,
                                    eps, rope ? rope->dimensions : 0, rope ? rope->traditional : false,
                                    fast::opt(rope ? std::optional<float>(rope->base) : std::nullopt),
                                    rope ? rope->scale : 1.f, scale, s.raw()));
  return out;
}

/// The fused decode step with the position read by the kernel from `position` (device int32): capturable
/// with cudaStreamBeginCapture and replayable per token.  `out` is caller-owned (static under a graph).
inline void attention_decode_fused_dynamic(Array& out, const Array& queries, const Array& keys, const Array& values,
                                           KVCache& cache, const nn::Rope* rope, float scale, const int32_t* position,
                                           const nn::RmsNorm* q_norm = nullptr, const nn::RmsNorm* k_norm = nullptr,
                                           Stream s = {}) {
  const float eps = q_norm ? q_norm->eps : (k_norm ? k_norm->eps : 0.f);
  check(omx_attn_decode_fused_dynamic(out.desc(), queries.desc(), keys.desc(), values.desc(), cache.raw(),
                                      q_norm ? q_norm->weight.desc() : nullptr,
                                      k_norm ? k_norm->weight.desc() : nullptr, eps, rope ? rope->dimensions : 0,
                                      rope ? rope->traditional : false,
                                      fast::opt(rope ? std::optional<float>(rope->base) : std::nullopt),
                                      rope ? rope->scale : 1.f, scale, position, s.raw()));
}
/// kv-head-sharded decode step of one rank with the data + flag exchange (see omx_attn_decode_fused_sharded_ll):
/// queries / keys / values carry the LOCAL heads, `out_full` is this rank's private [B,Hq_total,1,D] output and is
/// complete when the launch completes; `group` = every rank's staging buffer as mapped here + the local step counter.
inline void attention_decode_fused_sharded(Array& out_full, const Array& queries, const Array& keys,
                                           const Array& values, KVCache& cache, const nn::Rope* rope, float scale,
                                           const omx_ll_group& group, Stream s = {}) {
  const auto qs = queries.shape();
  if (qs.size() != 4 || qs[2] != 1) throw Exception("attention_decode_fused_sharded: queries must be [B, H, 1, D]");
  check(omx_attn_decode_fused_sharded_ll(out_full.desc(), queries.desc(), keys.desc(), values.desc(), cache.raw(),
                                         rope ? rope->dimensions : 0, rope ? rope->traditional : false,
                                         fast::opt(rope ? std::optional<float>(rope->base) : std::nullopt),
                                         rope ? rope->scale : 1.f, nullptr, scale, &group,
                                         (int)(group.rank * qs[1]), s.raw()));
}
/// Sequence-sharded decode step of one rank (see omx_attn_decode_seqshard): `partial` = this rank's float32
/// [world,B,Hq,D+2] buffer, `position` = GLOBAL position of the new token, `append` on the owning rank only.
inline void attention_decode_seqshard(Array& partial, const Array& queries, const Array* keys, const Array* values,
                                      KVCache& cache, const nn::Rope* rope, int position, bool append, float scale,
                                      const omx_peer_group* peers = nullptr, Stream s = {}) {
  check(omx_attn_decode_seqshard(partial.desc(), queries.desc(), keys ? keys->desc() : nullptr,
                                 values ? values->desc() : nullptr, cache.raw(), rope ? rope->dimensions : 0,
                                 rope ? rope->traditional : false,
                                 fast::opt(rope ? std::optional<float>(rope->base) : std::nullopt),
                                 rope ? rope->scale : 1.f, position, append, scale, peers, s.raw()));
}
/// Waits for the `expected`-th arrival of every rank (peers may be null: no wait), then merges the partials.
inline void seqshard_merge(Array& out, const Array& partial, const omx_peer_group* peers, uint32_t expected,
                           Stream s = {}) {
  check(omx_seqshard_merge(out.desc(), partial.desc(), peers, expected, s.raw()));
}
inline void device_counter_add(int32_t* counter, int delta, Stream s = {}) {
  check(omx_device_counter_add(counter, delta, s.raw()));
}

/// Attention::forward for L >= 1 new tokens with the minimum of memory passes (see omx_attn_prefill_fused).
inline Array attention_prefill_fused(const Array& queries, const Array& keys, const Array& values, KVCache& cache,
                                     const nn::Rope* rope, float scale, SdpaMask mask = {},
                                     const nn::RmsNorm* q_norm = nullptr, const nn::RmsNorm* k_norm = nullptr,
                                     Stream s = {}) {
  const auto qs = queries.shape();
  if (qs.size() != 4) throw Exception("attention_prefill_fused: queries must be [B, H, L, D]");
  Array out = Array::empty({qs[0], qs[1], qs[2], values.shape()[3]}, queries.dtype());
  const float eps = q_norm ? q_norm->eps : (k_norm ? k_norm->eps : 0.f);
  // the callers' rule (qwen3-mlx/src/model.rs:203-207): no mask and L > 1 means Causal
  if (mask.kind == SdpaMask::None && qs[2] > 1) mask = SdpaMask::causal();
  const auto m = mask.lower();
  check(omx_attn_prefill_fused(out.desc(), queries.desc(), keys.desc(), values.desc(), cache.raw(),
                               q_norm ? q_norm->weight.desc() : nullptr, k_norm ? k_norm->weight.desc() : nullptr, eps,
                               rope ? rope->dimensions : 0, rope ? rope->traditional : false,
                               fast::opt(rope ? std::optional<float>(rope->base) : std::nullopt),
                               rope ? rope->scale : 1.f, nullptr, scale, m.mode(), m.ptr(), nullptr, nullptr, s.raw()));
  return out;
}

}  // namespace utils
}  // namespace omx
