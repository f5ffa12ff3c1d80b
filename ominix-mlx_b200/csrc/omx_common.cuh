// omx_common.cuh -- shared host/device helpers for libomx_attn (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>

#include "../../include/omx_attn.h"

namespace omx {

// ---- errors -----------------------------------------------------------------
struct Error : public std::runtime_error {
  using std::runtime_error::runtime_error;
};

#define OMX_CHECK(cond, ...)                                        \
  do {                                                              \
    if (!(cond)) {                                                  \
      char _buf[512];                                               \
      snprintf(_buf, sizeof(_buf), __VA_ARGS__);                    \
      throw ::omx::Error(_buf);                                     \
    }                                                               \
  } while (0)

#define OMX_CUDA(expr)                                                          \
  do {                                                                          \
    cudaError_t _e = (expr);                                                    \
    if (_e != cudaSuccess) {                                                    \
      char _buf[512];                                                           \
      snprintf(_buf, sizeof(_buf), "CUDA error %s at %s:%d: %s", #expr,         \
               __FILE__, __LINE__, cudaGetErrorString(_e));                     \
      throw ::omx::Error(_buf);                                                 \
    }                                                                           \
  } while (0)

void note_launch(const char* kernel_family);  // omx_api.cu: thread-local launch accounting
void count_launch();

// ---- dtype ------------------------------------------------------------------
inline size_t dtype_size(int dt) {
  switch (dt) {
    case OMX_FLOAT32: return 4;
    case OMX_INT32: return 4;
    case OMX_FLOAT16: return 2;
    case OMX_BFLOAT16: return 2;
    case OMX_BOOL: return 1;
    default: return 0;
  }
}
inline bool is_float_dtype(int dt) {
  return dt == OMX_FLOAT32 || dt == OMX_FLOAT16 || dt == OMX_BFLOAT16;
}
inline const char* dtype_name(int dt) {
  switch (dt) {
    case OMX_FLOAT32: return "float32";
    case OMX_FLOAT16: return "float16";
    case OMX_BFLOAT16: return "bfloat16";
    case OMX_BOOL: return "bool";
    case OMX_INT32: return "int32";
    default: return "unsupported";
  }
}

// 4-D strided view used by the kernels ([B,H,L,D] order), element strides.
struct View4 {
  void* p;
  int64_t n[4];
  int64_t s[4];
};

inline View4 view4(const omx_array* a) {
  View4 v;
  v.p = a->data;
  for (int i = 0; i < 4; ++i) {
    v.n[i] = a->shape[i];
    v.s[i] = a->strides[i];
  }
  return v;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ---- device-side scalar conversion with explicit rounding ---------------------
template <typename T>
struct Num;
template <>
struct Num<float> {
  static __device__ __forceinline__ float to_f(float x) { return x; }
  static __device__ __forceinline__ float from_f(float x) { return x; }
  static __device__ __forceinline__ float lowest() { return -3.4028234663852886e38f; }
};
template <>
struct Num<__nv_bfloat16> {
  static __device__ __forceinline__ float to_f(__nv_bfloat16 x) { return __bfloat162float(x); }
  static __device__ __forceinline__ __nv_bfloat16 from_f(float x) { return __float2bfloat16_rn(x); }
  static __device__ __forceinline__ float lowest() { return -3.3895313892515355e38f; }
};
template <>
struct Num<__half> {
  static __device__ __forceinline__ float to_f(__half x) { return __half2float(x); }
  static __device__ __forceinline__ __half from_f(float x) { return __float2half_rn(x); }
  static __device__ __forceinline__ float lowest() { return -65504.0f; }
};

// value after a store to an array of type T (round-trip)
template <typename T>
__device__ __forceinline__ float rnd(float x) {
  return Num<T>::to_f(Num<T>::from_f(x));
}
template <>
__device__ __forceinline__ float rnd<float>(float x) {
  return x;
}

// One rope pair with the reference's per-primitive rounding (MLX CPU fallback graph):
//   o1 = T(T(x1*c) - T(x2*s)),  o2 = T(T(x1*s) + T(x2*c)),  c/s already rounded to T.
// __fmul_rn/__fadd_rn forbid FMA contraction (MLX runs separate multiply/subtract kernels).
template <typename T>
__device__ __forceinline__ void rope_pair(float x1, float x2, float c, float s, float& o1,
                                          float& o2) {
  float a = rnd<T>(__fmul_rn(x1, c));
  float b = rnd<T>(__fmul_rn(x2, s));
  float e = rnd<T>(__fmul_rn(x1, s));
  float f = rnd<T>(__fmul_rn(x2, c));
  o1 = rnd<T>(__fsub_rn(a, b));
  o2 = rnd<T>(__fadd_rn(e, f));
}

// rms_norm with the reference's per-primitive rounding (MLX CPU fallback graph, see norm.cu):
// r = 1 / sqrt(sum_d(x_d^2, left to right) * (1/D) + eps), all IEEE-rounded f32 operations.
template <typename T>
__device__ __forceinline__ float rms_rsqrt_row(const T* x, int64_t stride, int D, float eps, float inv_n) {
  float acc = 0.f;
  for (int d = 0; d < D; ++d) {
    const float v = Num<T>::to_f(x[d * stride]);
    acc = __fadd_rn(acc, __fmul_rn(v, v));
  }
  return __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(__fmul_rn(acc, inv_n), eps)));
}
// y = T(x * r); out = T(w * y)  (value returned as float, already rounded to T)
template <typename T>
__device__ __forceinline__ float rms_apply(float x, float r, float w, bool has_w) {
  const float y = rnd<T>(__fmul_rn(x, r));
  return has_w ? rnd<T>(__fmul_rn(w, y)) : y;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

constexpr float kLog2e = 1.4426950408889634f;

}  // namespace omx
