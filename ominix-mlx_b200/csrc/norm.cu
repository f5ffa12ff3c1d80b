// norm.cu -- fast::rms_norm over the last axis (the op right before the attention path: the
// per-head q_norm / k_norm of Qwen3-family crates, qwen3-mlx/src/model.rs:172-181).
//
// Replaces mlx_fast_rms_norm (mlx-c/mlx/c/fast.h:163-168, bound at mlx-rs/src/fast.rs:163-180).
// Numerics contract = the MLX CPU fallback graph, op for op, so that a KV cache filled through
// rms_norm -> rope -> append stays bit-identical to the reference's:
//   m = (sum_d float(x_d)^2, left to right in f32) * f32(1/D);  r = 1 / sqrt(m + eps)  (IEEE)
//   y_d = T(float(x_d) * r);  out_d = T(w_d * y_d)
// One thread owns one row: the left-to-right f32 sum is the definition of the result, so the row is
// not split; rows are independent, loads / stores are 128-bit when the layout allows.
#include "omx_common.cuh"
#include "omx_internal.h"

namespace omx {

namespace {

struct NormParams {
  const void* x;
  const void* w;
  void* out;
  int64_t n[OMX_MAX_NDIM];   // leading shape (row index space), innermost first excluded
  int64_t xs[OMX_MAX_NDIM];  // x strides of the leading axes
  int64_t os[OMX_MAX_NDIM];
  int nlead;
  int64_t rows;
  int D;
  int64_t x_inner, o_inner, w_inner;
  float eps, inv_n;
};

template <typename T>
__global__ void rms_norm_rows_kernel(const NormParams p) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= p.rows) return;
  int64_t rem = r, xo = 0, oo = 0;
  for (int i = p.nlead - 1; i >= 0; --i) {
    const int64_t c = rem % p.n[i];
    rem /= p.n[i];
    xo += c * p.xs[i];
    oo += c * p.os[i];
  }
  const T* x = (const T*)p.x + xo;
  T* o = (T*)p.out + oo;
  const T* w = (const T*)p.w;
  const float rs = rms_rsqrt_row<T>(x, p.x_inner, p.D, p.eps, p.inv_n);
  for (int d = 0; d < p.D; ++d) {
    const float v = Num<T>::to_f(x[d * p.x_inner]);
    o[d * p.o_inner] = Num<T>::from_f(rms_apply<T>(v, rs, w ? Num<T>::to_f(w[d * p.w_inner]) : 0.f, w != nullptr));
  }
}

// contiguous, 16-byte aligned rows of any length that is a multiple of 16 bytes: 128-bit loads, several in flight
// (the element-wise kernel above waits out one memory round trip per element: 256-wide rows -- Qwen3.5's q_norm /
// k_norm -- took ~130 us per call on a handful of rows).  Two passes: sum, then reload (L1 / L2) and scale.
template <typename T>
__global__ void rms_norm_chunk_kernel(const NormParams p, int w_vec) {
  constexpr int VE = 16 / sizeof(T);
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= p.rows) return;
  int64_t rem = r, xo = 0, oo = 0;
  for (int i = p.nlead - 1; i >= 0; --i) {
    const int64_t c = rem % p.n[i];
    rem /= p.n[i];
    xo += c * p.xs[i];
    oo += c * p.os[i];
  }
  const uint4* xv = reinterpret_cast<const uint4*>((const T*)p.x + xo);
  uint4* ov = reinterpret_cast<uint4*>((T*)p.out + oo);
  const int nv = p.D / VE;
  float acc = 0.f;
#pragma unroll 8
  for (int i = 0; i < nv; ++i) {
    const uint4 c = xv[i];
    const T* e = reinterpret_cast<const T*>(&c);
#pragma unroll
    for (int j = 0; j < VE; ++j) {
      const float v = Num<T>::to_f(e[j]);
      acc = __fadd_rn(acc, __fmul_rn(v, v));
    }
  }
  const float rs = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(__fmul_rn(acc, p.inv_n), p.eps)));
  const T* w = (const T*)p.w;
#pragma unroll 4
  for (int i = 0; i < nv; ++i) {
    uint4 c = xv[i];
    T* e = reinterpret_cast<T*>(&c);
    uint4 wc = make_uint4(0, 0, 0, 0);
    if (w && w_vec) wc = reinterpret_cast<const uint4*>(w)[i];
    const T* we = reinterpret_cast<const T*>(&wc);
#pragma unroll
    for (int j = 0; j < VE; ++j) {
      const float wj = !w ? 0.f : (w_vec ? Num<T>::to_f(we[j]) : Num<T>::to_f(w[(i * VE + j) * p.w_inner]));
      e[j] = Num<T>::from_f(rms_apply<T>(Num<T>::to_f(e[j]), rs, wj, w != nullptr));
    }
    ov[i] = c;
  }
}

// contiguous 16-bit / 32-bit rows of a compile-time length: whole row in registers, 128-bit I/O
template <typename T, int D>
__global__ void rms_norm_vec_kernel(const NormParams p) {
  constexpr int VE = 16 / sizeof(T);
  constexpr int NV = D / VE;
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= p.rows) return;
  int64_t rem = r, xo = 0, oo = 0;
  for (int i = p.nlead - 1; i >= 0; --i) {
    const int64_t c = rem % p.n[i];
    rem /= p.n[i];
    xo += c * p.xs[i];
    oo += c * p.os[i];
  }
  union Row { uint4 v[NV]; T t[D]; };
  Row row;
  const uint4* xv = reinterpret_cast<const uint4*>((const T*)p.x + xo);
#pragma unroll
  for (int i = 0; i < NV; ++i) row.v[i] = xv[i];
  float acc = 0.f;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    const float v = Num<T>::to_f(row.t[d]);
    acc = __fadd_rn(acc, __fmul_rn(v, v));
  }
  const float rs = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(__fmul_rn(acc, p.inv_n), p.eps)));
  const T* w = (const T*)p.w;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    const float v = Num<T>::to_f(row.t[d]);
    row.t[d] = Num<T>::from_f(rms_apply<T>(v, rs, w ? Num<T>::to_f(w[d * p.w_inner]) : 0.f, w != nullptr));
  }
  uint4* ov = reinterpret_cast<uint4*>((T*)p.out + oo);
#pragma unroll
  for (int i = 0; i < NV; ++i) ov[i] = row.v[i];
}

template <typename T>
void launch(const NormParams& p, bool vec_ok, cudaStream_t s) {
  const int threads = 128;
  const unsigned blocks = (unsigned)((p.rows + threads - 1) / threads);
  if (vec_ok && p.D == 128) {
    rms_norm_vec_kernel<T, 128><<<blocks, threads, 0, s>>>(p);
  } else if (vec_ok && p.D == 64) {
    rms_norm_vec_kernel<T, 64><<<blocks, threads, 0, s>>>(p);
  } else if (vec_ok && p.D % (16 / (int)sizeof(T)) == 0) {
    const int w_vec = p.w && p.w_inner == 1 && aligned16(p.w) ? 1 : 0;
    const int th = 32;  // few rows per call in practice (decode): spread them over SMs
    rms_norm_chunk_kernel<T><<<(unsigned)((p.rows + th - 1) / th), th, 0, s>>>(p, w_vec);
  } else {
    rms_norm_rows_kernel<T><<<blocks, threads, 0, s>>>(p);
  }
}

}  // namespace

void rms_norm_forward(const omx_array* out, const omx_array* x, const omx_array* weight, float eps,
                      cudaStream_t stream) {
  OMX_CHECK(out && x, "[rms_norm] null array");
  OMX_CHECK(x->ndim >= 1 && x->ndim <= OMX_MAX_NDIM, "[rms_norm] Input must have at least 1 dimension but got input with "
                                                      "%d dimensions.", x->ndim);
  OMX_CHECK(is_float_dtype(x->dtype), "[rms_norm] Received unsupported type %s.", dtype_name(x->dtype));
  const int nd = x->ndim;
  const int64_t D = x->shape[nd - 1];
  const bool has_w = weight && weight->data;
  if (has_w) {
    OMX_CHECK(weight->ndim == 1, "[rms_norm] weight must have 1 dimension but has %d dimensions.", weight->ndim);
    OMX_CHECK(weight->shape[0] == D, "[rms_norm] weight must have the same size as the last dimension of x but has "
                                     "%lld elements.", (long long)weight->shape[0]);
    OMX_CHECK(weight->dtype == x->dtype, "[rms_norm] weight dtype %s differs from x dtype %s (the reference "
                                         "promotes; this boundary does not)", dtype_name(weight->dtype),
              dtype_name(x->dtype));
  }
  OMX_CHECK(out->ndim == nd && out->dtype == x->dtype, "[rms_norm] out must have x's rank and dtype");
  for (int i = 0; i < nd; ++i) OMX_CHECK(out->shape[i] == x->shape[i], "[rms_norm] out must have x's shape");
  NormParams p{};
  p.x = x->data;
  p.out = out->data;
  p.w = has_w ? weight->data : nullptr;
  p.nlead = nd - 1;
  p.rows = 1;
  bool vec_ok = x->strides[nd - 1] == 1 && out->strides[nd - 1] == 1 && aligned16(x->data) && aligned16(out->data);
  const int64_t ve = (int64_t)(16 / dtype_size(x->dtype));
  for (int i = 0; i < nd - 1; ++i) {
    p.n[i] = x->shape[i];
    p.xs[i] = x->strides[i];
    p.os[i] = out->strides[i];
    p.rows *= x->shape[i];
    if (x->shape[i] > 1 && (x->strides[i] % ve || out->strides[i] % ve)) vec_ok = false;
  }
  p.D = (int)D;
  p.x_inner = x->strides[nd - 1];
  p.o_inner = out->strides[nd - 1];
  p.w_inner = has_w ? weight->strides[0] : 0;
  p.eps = eps;
  p.inv_n = 1.0f / (float)D;
  if (p.rows == 0 || D == 0) return;
  note_launch("rms_norm");
  switch (x->dtype) {
    case OMX_FLOAT32: launch<float>(p, vec_ok, stream); break;
    case OMX_BFLOAT16: launch<__nv_bfloat16>(p, vec_ok, stream); break;
    default: launch<__half>(p, vec_ok, stream); break;
  }
  count_launch();
  OMX_CUDA(cudaGetLastError());
}

}  // namespace omx
