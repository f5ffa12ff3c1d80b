// rope.cu -- RoPE for the B200 attention path.
//
// Replaces mlx_fast_rope / mlx_fast_rope_dynamic (mlx-c/mlx/c/fast.h:169-188, bound at
// mlx-rs/src/fast.rs:31-45) and the table-driven DiT rope of the image crates
// (flux-klein-mlx/src/klein_model.rs:124-162, zimage-mlx/src/zimage_model.rs:208-235).
//
// Numerics contract (bit parity with the MLX CPU backend's fallback graph):
//   * theta = ((t + offset) * scale) * inv_freq with inv_freq = expf(-i * (logf(base)/half))
//     (or 1/freqs), cos/sin from the HOST libm -- the table is built on the CPU with the very
//     calls the reference's CPU path makes, then kept resident in HBM (a few MB at most).
//   * the rotation runs in x's dtype with a rounding after every multiply / add / subtract.
#include <cmath>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <vector>

#include "omx_common.cuh"
#include "omx_internal.h"

namespace omx {

// ---------------------------------------------------------------- table cache
namespace {

struct RopeTable {
  int device = 0;
  int dims = 0;
  bool has_base = false;
  float base = 0.f;
  float scale = 1.f;
  std::vector<float> freqs;  // host copy when given
  int n_pos = 0;
  float* d_cos = nullptr;  // [n_pos, half]
  float* d_sin = nullptr;
  // the same rows rounded once to bf16 ([0]) / f16 ([1]): what the 16-bit kernels multiply with (prologue.cu)
  void* d_cos16[2] = {nullptr, nullptr};
  void* d_sin16[2] = {nullptr, nullptr};
};

std::mutex g_tbl_mu;
std::vector<std::unique_ptr<RopeTable>> g_tables;

void fill_rows(const RopeTable& t, int p0, int p1, float* c, float* s) {
  const int half = t.dims / 2;
  std::vector<float> inv(half);
  if (!t.freqs.empty()) {
    for (int i = 0; i < half; ++i) inv[i] = 1.0f / t.freqs[i];
  } else {
    const float step = logf(t.base) / (float)half;
    for (int i = 0; i < half; ++i) inv[i] = expf((float)(-i) * step);
  }
  for (int p = p0; p < p1; ++p) {
    // (arange(T) + offset) * scale: float(t) + float(offset) is exact below 2^24
    const float pos = (float)p * t.scale;
    for (int i = 0; i < half; ++i) {
      const float th = pos * inv[i];
      c[(size_t)(p - p0) * half + i] = cosf(th);
      s[(size_t)(p - p0) * half + i] = sinf(th);
    }
  }
}

}  // namespace

RopeTableRef get_rope_table(int dims, bool has_base, float base, float scale,
                            const float* freqs_host, int need_positions, cudaStream_t stream) {
  OMX_CHECK(dims > 0 && dims % 2 == 0, "[rope] dims must be positive and even, got %d", dims);
  OMX_CHECK(need_positions < (1 << 24), "[rope] positions beyond 2^24 are not supported");
  const int half = dims / 2;
  int dev = 0;
  OMX_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(g_tbl_mu);
  RopeTable* t = nullptr;
  for (auto& u : g_tables) {
    if (u->device != dev || u->dims != dims || u->scale != scale) continue;
    if (freqs_host) {
      if (u->freqs.size() == (size_t)half &&
          memcmp(u->freqs.data(), freqs_host, sizeof(float) * half) == 0) {
        t = u.get();
        break;
      }
    } else if (u->freqs.empty() && u->has_base && u->base == base) {
      t = u.get();
      break;
    }
  }
  if (!t) {
    g_tables.emplace_back(new RopeTable());
    t = g_tables.back().get();
    t->device = dev;
    t->dims = dims;
    t->has_base = has_base;
    t->base = base;
    t->scale = scale;
    if (freqs_host) t->freqs.assign(freqs_host, freqs_host + half);
  }
  if (need_positions > t->n_pos) {
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    OMX_CUDA(cudaStreamIsCapturing(stream, &cap));
    OMX_CHECK(cap == cudaStreamCaptureStatusNone, "[rope] the position table for these parameters must be built "
              "before stream capture: run the same call once eagerly first");
    int n = std::max(need_positions, std::max(2 * t->n_pos, 4096));
    n = (n + 1023) / 1024 * 1024;
    std::vector<float> c((size_t)n * half), s((size_t)n * half);
    fill_rows(*t, 0, n, c.data(), s.data());
    float *dc = nullptr, *ds = nullptr;
    OMX_CUDA(cudaMalloc(&dc, sizeof(float) * c.size()));
    OMX_CUDA(cudaMalloc(&ds, sizeof(float) * s.size()));
    OMX_CUDA(cudaMemcpyAsync(dc, c.data(), sizeof(float) * c.size(), cudaMemcpyHostToDevice, stream));
    OMX_CUDA(cudaMemcpyAsync(ds, s.data(), sizeof(float) * s.size(), cudaMemcpyHostToDevice, stream));
    // 16-bit copies, rounded on the host with the same round-to-nearest-even the kernels' conversions use
    std::vector<uint16_t> h16(c.size());
    void* d16[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
    for (int ty = 0; ty < 2; ++ty) {
      for (int cs = 0; cs < 2; ++cs) {
        const std::vector<float>& src = cs ? s : c;
        for (size_t i = 0; i < src.size(); ++i) {
          if (ty == 0) {
            const __nv_bfloat16 v = __float2bfloat16_rn(src[i]);
            memcpy(&h16[i], &v, 2);
          } else {
            const __half v = __float2half_rn(src[i]);
            memcpy(&h16[i], &v, 2);
          }
        }
        OMX_CUDA(cudaMalloc(&d16[ty][cs], sizeof(uint16_t) * h16.size()));
        // (pageable source: the call returns once the staging copy is done, so h16 can be reused)
        OMX_CUDA(cudaMemcpyAsync(d16[ty][cs], h16.data(), sizeof(uint16_t) * h16.size(), cudaMemcpyHostToDevice, stream));
      }
    }
    // Rare (creation / doubling): make the table visible to every stream before first use.
    OMX_CUDA(cudaStreamSynchronize(stream));
    for (int ty = 0; ty < 2; ++ty) {
      t->d_cos16[ty] = d16[ty][0];
      t->d_sin16[ty] = d16[ty][1];
    }
    // Old buffers may still be read by kernels in flight on other streams: retire, never free.
    t->d_cos = dc;
    t->d_sin = ds;
    t->n_pos = n;
  }
  RopeTableRef r;
  r.cos = t->d_cos;
  r.sin = t->d_sin;
  r.half = half;
  r.n_pos = t->n_pos;
  for (int ty = 0; ty < 2; ++ty) {
    r.cos16[ty] = t->d_cos16[ty];
    r.sin16[ty] = t->d_sin16[ty];
  }
  return r;
}

// -------------------------------------------------------------------- kernels
namespace {

struct RopeParams {
  const void* x;
  void* out;
  int64_t xs[4], os[4];  // element strides of the [B,N,T,D] views
  int B, N, T, D, dims;
  const float* cos;  // [n_pos, half]
  const float* sin;
  int offset;
  const int32_t* offset_dev;  // overrides offset when non-null
  int n_pos;
};

// Scalar kernel: one thread per (row, pair) / per tail element.  Any strides.
template <typename T, bool TRAD>
__global__ void rope_scalar_kernel(RopeParams p) {
  const int half = p.dims / 2;
  const int per_row = half + (p.D - p.dims);
  const int64_t total = (int64_t)p.B * p.N * p.T * per_row;
  const int off = p.offset_dev ? *p.offset_dev : p.offset;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int u = (int)(idx % per_row);
    int64_t r = idx / per_row;
    const int t = (int)(r % p.T);
    r /= p.T;
    const int n = (int)(r % p.N);
    const int b = (int)(r / p.N);
    const T* x = (const T*)p.x + b * p.xs[0] + n * p.xs[1] + t * p.xs[2];
    T* o = (T*)p.out + b * p.os[0] + n * p.os[1] + t * p.os[2];
    if (u < half) {
      const int i1 = TRAD ? 2 * u : u;
      const int i2 = TRAD ? 2 * u + 1 : u + half;
      const int pos = min(off + t, p.n_pos - 1);
      const float c = rnd<T>(p.cos[(size_t)pos * half + u]);
      const float s = rnd<T>(p.sin[(size_t)pos * half + u]);
      float o1, o2;
      rope_pair<T>(Num<T>::to_f(x[i1 * p.xs[3]]), Num<T>::to_f(x[i2 * p.xs[3]]), c, s, o1, o2);
      o[i1 * p.os[3]] = Num<T>::from_f(o1);
      o[i2 * p.os[3]] = Num<T>::from_f(o2);
    } else {
      const int d = p.dims + (u - half);
      o[d * p.os[3]] = x[d * p.xs[3]];
    }
  }
}

template <typename T>
struct Vec16 {
  static constexpr int N = 16 / sizeof(T);
  union {
    uint4 raw;
    T v[N];
  };
};

// Vector kernel: last axis contiguous, 16-byte accesses, dims == D.
// non-traditional: a thread owns pairs [i, i+V) -> two 16 B loads / stores (x1 block, x2 block)
// traditional    : a thread owns V/2... kept simple: V elements = V/2 adjacent pairs per access,
//                  two accesses per thread so both variants move 32 B per thread.
template <typename T, bool TRAD>
__global__ void rope_vec_kernel(RopeParams p) {
  constexpr int V = Vec16<T>::N;
  const int half = p.dims / 2;
  const int per_row = half / V;  // units of V pairs
  const int64_t total = (int64_t)p.B * p.N * p.T * per_row;
  const int off = p.offset_dev ? *p.offset_dev : p.offset;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int u = (int)(idx % per_row);
    int64_t r = idx / per_row;
    const int t = (int)(r % p.T);
    r /= p.T;
    const int n = (int)(r % p.N);
    const int b = (int)(r / p.N);
    const T* x = (const T*)p.x + b * p.xs[0] + n * p.xs[1] + t * p.xs[2];
    T* o = (T*)p.out + b * p.os[0] + n * p.os[1] + t * p.os[2];
    const int pos = min(off + t, p.n_pos - 1);
    const float* cr = p.cos + (size_t)pos * half + u * V;
    const float* sr = p.sin + (size_t)pos * half + u * V;
    Vec16<T> a, bb, oa, ob;
    if (!TRAD) {
      a.raw = *reinterpret_cast<const uint4*>(x + u * V);
      bb.raw = *reinterpret_cast<const uint4*>(x + half + u * V);
#pragma unroll
      for (int j = 0; j < V; ++j) {
        float o1, o2;
        rope_pair<T>(Num<T>::to_f(a.v[j]), Num<T>::to_f(bb.v[j]), rnd<T>(cr[j]), rnd<T>(sr[j]), o1, o2);
        oa.v[j] = Num<T>::from_f(o1);
        ob.v[j] = Num<T>::from_f(o2);
      }
      *reinterpret_cast<uint4*>(o + u * V) = oa.raw;
      *reinterpret_cast<uint4*>(o + half + u * V) = ob.raw;
    } else {
      a.raw = *reinterpret_cast<const uint4*>(x + 2 * u * V);
      bb.raw = *reinterpret_cast<const uint4*>(x + 2 * u * V + V);
#pragma unroll
      for (int j = 0; j < V / 2; ++j) {
        float o1, o2;
        rope_pair<T>(Num<T>::to_f(a.v[2 * j]), Num<T>::to_f(a.v[2 * j + 1]), rnd<T>(cr[j]), rnd<T>(sr[j]), o1, o2);
        oa.v[2 * j] = Num<T>::from_f(o1);
        oa.v[2 * j + 1] = Num<T>::from_f(o2);
        rope_pair<T>(Num<T>::to_f(bb.v[2 * j]), Num<T>::to_f(bb.v[2 * j + 1]), rnd<T>(cr[V / 2 + j]),
                     rnd<T>(sr[V / 2 + j]), o1, o2);
        ob.v[2 * j] = Num<T>::from_f(o1);
        ob.v[2 * j + 1] = Num<T>::from_f(o2);
      }
      *reinterpret_cast<uint4*>(o + 2 * u * V) = oa.raw;
      *reinterpret_cast<uint4*>(o + 2 * u * V + V) = ob.raw;
    }
  }
}

struct DitRopeParams {
  const void *x, *cos, *sin;
  void* out;
  int64_t xs[4], os[4], cs[3], ss[3];
  int B, S, H, D;
};

// DiT: x [B,S,H,D], cos/sin [B,S,D/2] in T; adjacent pairs.
template <typename T>
__global__ void dit_rope_kernel(DitRopeParams p) {
  const int half = p.D / 2;
  const int64_t total = (int64_t)p.B * p.S * p.H * half;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(idx % half);
    int64_t r = idx / half;
    const int h = (int)(r % p.H);
    r /= p.H;
    const int s = (int)(r % p.S);
    const int b = (int)(r / p.S);
    const T* x = (const T*)p.x + b * p.xs[0] + s * p.xs[1] + h * p.xs[2];
    T* o = (T*)p.out + b * p.os[0] + s * p.os[1] + h * p.os[2];
    const float c = Num<T>::to_f(((const T*)p.cos)[b * p.cs[0] + s * p.cs[1] + i * p.cs[2]]);
    const float sn = Num<T>::to_f(((const T*)p.sin)[b * p.ss[0] + s * p.ss[1] + i * p.ss[2]]);
    float o1, o2;
    rope_pair<T>(Num<T>::to_f(x[(2 * i) * p.xs[3]]), Num<T>::to_f(x[(2 * i + 1) * p.xs[3]]), c, sn, o1, o2);
    o[(2 * i) * p.os[3]] = Num<T>::from_f(o1);
    o[(2 * i + 1) * p.os[3]] = Num<T>::from_f(o2);
  }
}

template <typename T>
void launch_rope(const RopeParams& p, bool traditional, bool vec_ok, cudaStream_t s) {
  const int half = p.dims / 2;
  constexpr int V = 16 / sizeof(T);
  const int threads = 256;
  if (vec_ok) {
    const int64_t total = (int64_t)p.B * p.N * p.T * (half / V);
    const int blocks = (int)std::min<int64_t>((total + threads - 1) / threads, 148 * 16);
    if (traditional) rope_vec_kernel<T, true><<<blocks, threads, 0, s>>>(p);
    else rope_vec_kernel<T, false><<<blocks, threads, 0, s>>>(p);
  } else {
    const int64_t total = (int64_t)p.B * p.N * p.T * (half + p.D - p.dims);
    const int blocks = (int)std::min<int64_t>((total + threads - 1) / threads, 148 * 16);
    if (traditional) rope_scalar_kernel<T, true><<<blocks, threads, 0, s>>>(p);
    else rope_scalar_kernel<T, false><<<blocks, threads, 0, s>>>(p);
  }
  count_launch();
  OMX_CUDA(cudaGetLastError());
}

// Collapse x [..., T, D] (ndim >= 3) into [B, N, T, D] strides the way the fallback graph
// does (ndim 3: N = 1; ndim > 4: flatten axes 1..ndim-3).
void collapse(const omx_array* a, int64_t shape[4], int64_t strides[4], const char* what) {
  const int nd = a->ndim;
  shape[0] = a->shape[0];
  strides[0] = a->strides[0];
  shape[2] = a->shape[nd - 2];
  strides[2] = a->strides[nd - 2];
  shape[3] = a->shape[nd - 1];
  strides[3] = a->strides[nd - 1];
  if (nd == 3) {
    shape[1] = 1;
    strides[1] = 0;
    return;
  }
  int64_t n = 1;
  for (int i = 1; i <= nd - 3; ++i) n *= a->shape[i];
  for (int i = 1; i < nd - 3; ++i) {
    OMX_CHECK(a->shape[i + 1] == 1 || a->shape[i] == 1 ||
                  a->strides[i] == a->strides[i + 1] * a->shape[i + 1],
              "[rope] %s: axes 1..%d of a %d-d array must be collapsible", what, nd - 3, nd);
  }
  shape[1] = n;
  strides[1] = a->strides[nd - 3];
}

}  // namespace

void rope_forward(const omx_array* out, const omx_array* x, int dims, bool traditional,
                  omx_optional_float base, float scale, int offset, const omx_array* offset_arr,
                  int max_position, const omx_array* freqs, cudaStream_t stream) {
  OMX_CHECK(x && out, "[rope] null array");
  OMX_CHECK(x->ndim >= 3 && x->ndim <= OMX_MAX_NDIM,
            "[rope] Input must have at least 3 dimensions but got input with %d dimensions.", x->ndim);
  OMX_CHECK(is_float_dtype(x->dtype), "[rope] Input must be a floating type but got %s.",
            dtype_name(x->dtype));
  OMX_CHECK(out->dtype == x->dtype && out->ndim == x->ndim, "[rope] out must match x");
  for (int i = 0; i < x->ndim; ++i)
    OMX_CHECK(out->shape[i] == x->shape[i], "[rope] out shape must match x");
  OMX_CHECK(base.has_value != (freqs != nullptr && freqs->data != nullptr),
            "[rope] Only one of base or freqs can have a value.");
  const int D = (int)x->shape[x->ndim - 1];
  OMX_CHECK(dims > 0 && dims % 2 == 0 && dims <= D,
            "[rope] dims must be even and in (0, %d], got %d", D, dims);
  const int half = dims / 2;
  std::vector<float> fh;
  if (!base.has_value) {
    OMX_CHECK(freqs->ndim == 1 && freqs->shape[0] == half && freqs->dtype == OMX_FLOAT32,
              "[rope] freqs must be a float32 vector of length dims/2 = %d", half);
    // bit parity needs the host libm: fetch the (tiny) vector; this synchronises the stream.
    fh.resize(half);
    if (freqs->strides[0] == 1) {
      OMX_CUDA(cudaMemcpyAsync(fh.data(), freqs->data, sizeof(float) * half, cudaMemcpyDeviceToHost, stream));
    } else {
      OMX_CUDA(cudaMemcpy2DAsync(fh.data(), sizeof(float), freqs->data, sizeof(float) * freqs->strides[0],
                                 sizeof(float), half, cudaMemcpyDeviceToHost, stream));
    }
    OMX_CUDA(cudaStreamSynchronize(stream));
  }
  RopeParams p;
  int64_t xs[4], xn[4], os[4], on[4];
  collapse(x, xn, xs, "x");
  collapse(out, on, os, "out");
  p.x = x->data;
  p.out = out->data;
  for (int i = 0; i < 4; ++i) {
    p.xs[i] = xs[i];
    p.os[i] = os[i];
  }
  p.B = (int)xn[0];
  p.N = (int)xn[1];
  p.T = (int)xn[2];
  p.D = D;
  p.dims = dims;
  p.offset = offset;
  p.offset_dev = nullptr;
  int need;
  if (offset_arr) {
    OMX_CHECK(offset_arr->dtype == OMX_INT32 && offset_arr->data, "[rope] offset must be an int32 device scalar");
    OMX_CHECK(max_position > 0, "[rope] max_position must be given with a device offset");
    p.offset_dev = (const int32_t*)offset_arr->data;
    need = max_position + p.T;
  } else {
    OMX_CHECK(offset >= 0, "[rope] negative offsets are not supported (got %d)", offset);
    need = offset + p.T;
  }
  if ((int64_t)p.B * p.N * p.T * D == 0) return;
  OMX_CHECK(x->data && out->data, "[rope] null data pointer");
  RopeTableRef tb = get_rope_table(dims, base.has_value, base.value, scale,
                                   fh.empty() ? nullptr : fh.data(), need, stream);
  p.cos = tb.cos;
  p.sin = tb.sin;
  p.n_pos = tb.n_pos;
  const size_t es = dtype_size(x->dtype);
  const int V = (int)(16 / es);
  bool vec_ok = dims == D && half % V == 0 && xs[3] == 1 && os[3] == 1 && aligned16(x->data) &&
                aligned16(out->data);
  for (int i = 0; i < 3; ++i) vec_ok = vec_ok && (xs[i] % V == 0) && (os[i] % V == 0);
  note_launch("rope");
  switch (x->dtype) {
    case OMX_FLOAT32: launch_rope<float>(p, traditional, vec_ok, stream); break;
    case OMX_BFLOAT16: launch_rope<__nv_bfloat16>(p, traditional, vec_ok, stream); break;
    default: launch_rope<__half>(p, traditional, vec_ok, stream); break;
  }
}

void dit_rope_forward(const omx_array* out, const omx_array* x, const omx_array* cs,
                      const omx_array* sn, cudaStream_t stream) {
  OMX_CHECK(x && out && cs && sn, "[dit_rope] null array");
  OMX_CHECK(x->ndim == 4 && out->ndim == 4, "[dit_rope] x must be [B,S,H,D]");
  OMX_CHECK(is_float_dtype(x->dtype) && out->dtype == x->dtype && cs->dtype == x->dtype &&
                sn->dtype == x->dtype,
            "[dit_rope] x, out, cos, sin must share one floating dtype");
  DitRopeParams p;
  p.x = x->data;
  p.out = out->data;
  p.cos = cs->data;
  p.sin = sn->data;
  p.B = (int)x->shape[0];
  p.S = (int)x->shape[1];
  p.H = (int)x->shape[2];
  p.D = (int)x->shape[3];
  OMX_CHECK(p.D % 2 == 0, "[dit_rope] head_dim must be even");
  for (int i = 0; i < 4; ++i) {
    OMX_CHECK(out->shape[i] == x->shape[i], "[dit_rope] out shape must match x");
    p.xs[i] = x->strides[i];
    p.os[i] = out->strides[i];
  }
  // accept [B,S,D/2] or [B,S,1,D/2]
  auto tbl = [&](const omx_array* a, int64_t st[3], const char* nm) {
    if (a->ndim == 3) {
      OMX_CHECK(a->shape[0] == p.B && a->shape[1] == p.S && a->shape[2] == p.D / 2,
                "[dit_rope] %s must be [B,S,D/2]", nm);
      st[0] = a->strides[0]; st[1] = a->strides[1]; st[2] = a->strides[2];
    } else {
      OMX_CHECK(a->ndim == 4 && a->shape[0] == p.B && a->shape[1] == p.S && a->shape[2] == 1 &&
                    a->shape[3] == p.D / 2,
                "[dit_rope] %s must be [B,S,D/2] or [B,S,1,D/2]", nm);
      st[0] = a->strides[0]; st[1] = a->strides[1]; st[2] = a->strides[3];
    }
  };
  tbl(cs, p.cs, "cos");
  tbl(sn, p.ss, "sin");
  const int64_t total = (int64_t)p.B * p.S * p.H * (p.D / 2);
  if (total == 0) return;
  note_launch("dit_rope");
  {  // whole rows in registers, 128-bit accesses (prologue.cu) when the layout allows
    auto bhsd = [](const omx_array& a) {
      omx_array t = a;
      t.shape[1] = a.shape[2]; t.shape[2] = a.shape[1];
      t.strides[1] = a.strides[2]; t.strides[2] = a.strides[1];
      return t;
    };
    PrologueCall pc;
    const omx_array xv = bhsd(*x);
    pc.seg[0].x = &xv;
    pc.seg[0].out = bhsd(*out);
    pc.seg[0].rope = true;
    pc.nseg = 1;
    pc.mode = 2;
    pc.dims = p.D;
    pc.tcos = cs;
    pc.tsin = sn;
    for (int j = 0; j < 3; ++j) {
      pc.tcs[j] = p.cs[j];
      pc.tss[j] = p.ss[j];
    }
    if (x->data != out->data && qkv_prologue(pc, stream)) return;
  }
  const int threads = 256;
  const int blocks = (int)std::min<int64_t>((total + threads - 1) / threads, 148 * 16);
  switch (x->dtype) {
    case OMX_FLOAT32: dit_rope_kernel<float><<<blocks, threads, 0, stream>>>(p); break;
    case OMX_BFLOAT16: dit_rope_kernel<__nv_bfloat16><<<blocks, threads, 0, stream>>>(p); break;
    default: dit_rope_kernel<__half><<<blocks, threads, 0, stream>>>(p); break;
  }
  count_launch();
  OMX_CUDA(cudaGetLastError());
}

}  // namespace omx
