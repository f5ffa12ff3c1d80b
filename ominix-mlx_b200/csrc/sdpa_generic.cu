// sdpa_generic.cu -- shape/stride/mask-complete attention on CUDA cores.
//
// The catch-all behind omx_fast_scaled_dot_product_attention (replaces
// mlx_fast_scaled_dot_product_attention, mlx-c/mlx/c/fast.h:189-198) for everything the
// specialised kernels (decode.cu, fmha_sm100.cu) do not take: float32, head dims other than
// 128, boolean / additive array masks (mlx-rs-core/src/utils.rs:134-153 builds the bool ones),
// arbitrary strides.  One warp owns one query row and streams the keys with an online
// softmax; scores, softmax and the PV accumulation are float32 (mlx-rs/src/fast.rs:116).
//
// Mask semantics follow the MLX fallback graph: bool / causal entries that are masked take
// finfo(dtype).min (so a fully masked row degrades to a uniform average, not NaN); additive
// masks are added to the scaled scores.
#include <algorithm>
#include <type_traits>

#include "omx_common.cuh"
#include "omx_internal.h"

namespace omx {

namespace {

struct GenParams {
  const void *q, *k, *v, *mask;
  void* out;
  int64_t qs[4], ks[4], vs[4], os[4], ms[4];
  int B, Hq, Hkv, Lq, Lk, D, Dv;
  float scale;
  int mask_mode;
  int mask_is_f32;  // additive mask stored as float32 (else: T)
  int out_is_f32;   // output stored as float32 (else: T)
};

template <typename T, int E>
__global__ void __launch_bounds__(128) sdpa_generic_kernel(GenParams p) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= (int64_t)p.B * p.Hq * p.Lq) return;
  const int i = (int)(row % p.Lq);
  const int h = (int)((row / p.Lq) % p.Hq);
  const int b = (int)(row / ((int64_t)p.Lq * p.Hq));
  const int hk = h / (p.Hq / p.Hkv);
  const T* q = (const T*)p.q + b * p.qs[0] + h * p.qs[1] + i * p.qs[2];
  const T* kb = (const T*)p.k + b * p.ks[0] + hk * p.ks[1];
  const T* vb = (const T*)p.v + b * p.vs[0] + hk * p.vs[1];
  const int64_t mrow = b * p.ms[0] + h * p.ms[1] + i * p.ms[2];

  float qr[E], acc[E];
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int d = lane + 32 * e;
    qr[e] = d < p.D ? Num<T>::to_f(q[d * p.qs[3]]) : 0.f;
    acc[e] = 0.f;
  }
  const int q_off = max(p.Lk - p.Lq, 0);
  const int jmax = p.mask_mode == MASK_CAUSAL ? min(p.Lk, q_off + i + 1) : p.Lk;
  const float fill = Num<T>::lowest();
  float m = -INFINITY, l = 0.f;

  for (int j0 = 0; j0 < jmax; j0 += 4) {
    float s[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = j0 + u;
      float a = 0.f;
      if (j < jmax) {
        const T* kr = kb + (int64_t)j * p.ks[2];
#pragma unroll
        for (int e = 0; e < E; ++e) {
          const int d = lane + 32 * e;
          if (d < p.D) a = fmaf(qr[e], Num<T>::to_f(kr[d * p.ks[3]]), a);
        }
      }
      s[u] = a;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) s[u] = warp_sum(s[u]);
    float tmax = -INFINITY;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = j0 + u;
      float x = s[u] * p.scale;
      if (j >= jmax) {
        x = -INFINITY;
      } else if (p.mask_mode == MASK_BOOL) {
        if (!((const uint8_t*)p.mask)[mrow + j * p.ms[3]]) x = fill;
      } else if (p.mask_mode == MASK_ADD) {
        const int64_t mi = mrow + j * p.ms[3];
        x += p.mask_is_f32 ? ((const float*)p.mask)[mi] : Num<T>::to_f(((const T*)p.mask)[mi]);
      }
      s[u] = x;
      tmax = fmaxf(tmax, x);
    }
    const float m_new = fmaxf(m, tmax);
    const float m_safe = (m_new == -INFINITY) ? 0.f : m_new;
    const float corr = exp2f((m - m_safe) * kLog2e);
    float psum = 0.f;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      s[u] = exp2f((s[u] - m_safe) * kLog2e);
      psum += s[u];
    }
    l = l * corr + psum;
#pragma unroll
    for (int e = 0; e < E; ++e) acc[e] *= corr;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = j0 + u;
      if (j < jmax) {
        const T* vr = vb + (int64_t)j * p.vs[2];
#pragma unroll
        for (int e = 0; e < E; ++e) {
          const int d = lane + 32 * e;
          if (d < p.Dv) acc[e] = fmaf(s[u], Num<T>::to_f(vr[d * p.vs[3]]), acc[e]);
        }
      }
    }
    m = m_new;
  }
  const float inv = 1.0f / l;
  const int64_t oo = b * p.os[0] + h * p.os[1] + i * p.os[2];
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int d = lane + 32 * e;
    if (d < p.Dv) {
      const float r = acc[e] * inv;
      if (p.out_is_f32) ((float*)p.out)[oo + d * p.os[3]] = r;
      else ((T*)p.out)[oo + d * p.os[3]] = Num<T>::from_f(r);
    }
  }
}

template <typename T>
void launch(const GenParams& p, cudaStream_t s) {
  const int64_t rows = (int64_t)p.B * p.Hq * p.Lq;
  if (rows == 0) return;
  const int wpb = 4;
  const unsigned blocks = (unsigned)((rows + wpb - 1) / wpb);
  const int need = (std::max(p.D, p.Dv) + 31) / 32;
  if (need <= 1) sdpa_generic_kernel<T, 1><<<blocks, wpb * 32, 0, s>>>(p);
  else if (need <= 2) sdpa_generic_kernel<T, 2><<<blocks, wpb * 32, 0, s>>>(p);
  else if (need <= 4) sdpa_generic_kernel<T, 4><<<blocks, wpb * 32, 0, s>>>(p);
  else if (need <= 8) sdpa_generic_kernel<T, 8><<<blocks, wpb * 32, 0, s>>>(p);
  else sdpa_generic_kernel<T, 20><<<blocks, wpb * 32, 0, s>>>(p);  // absorbed MLA: Dk = 576, Dv = 512
  count_launch();
  OMX_CUDA(cudaGetLastError());
}

// Rows whose array mask hides EVERY key.  The reference's CPU chain fills masked scores with
// finfo(dtype).min (bool masks) or adds ~-1e9 (the callers' additive masks), so such a row's
// softmax is uniform and its output is sum_k T(1/Lk) * V[k] -- an average over ALL Lk keys.  The
// tile-skipping kernels (fmha_sm100.cu, decode.cu) never visit masked keys; they flag those rows
// (dead[b][h][row] = 1) and this pass rewrites them.  One CTA per 128 query rows of one (b, h):
// a block without a flagged row exits after one byte load per thread.
template <typename T>
__global__ void __launch_bounds__(128)
masked_rows_fixup_kernel(const uint8_t* __restrict__ dead, const T* __restrict__ v, int64_t vs0, int64_t vs1,
                         int64_t vs2, int64_t vs3, void* out, int64_t os0, int64_t os1, int64_t os2, int64_t os3,
                         int Hq, int Hkv, int Lq, int Lk, int Dv, int out_is_f32) {
  __shared__ float mean[640];
  __shared__ uint8_t fl[128];
  const int tid = threadIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int row = blockIdx.x * 128 + tid;
  const int flag = row < Lq ? dead[((int64_t)b * Hq + h) * Lq + row] : 0;
  if (!__syncthreads_or(flag)) return;
  const T* vb = v + b * vs0 + (int64_t)(h / (Hq / Hkv)) * vs1;
  const float w = Num<T>::to_f(Num<T>::from_f(1.0f / (float)Lk));
  for (int d = tid; d < Dv; d += 128) {
    float acc = 0.f;
#pragma unroll 8
    for (int k = 0; k < Lk; ++k) acc = fmaf(w, Num<T>::to_f(vb[(int64_t)k * vs2 + d * vs3]), acc);
    mean[d] = acc;
  }
  fl[tid] = (uint8_t)flag;
  __syncthreads();
  for (int r = 0; r < 128; ++r) {
    if (!fl[r]) continue;
    const int64_t oo = b * os0 + h * os1 + (int64_t)(blockIdx.x * 128 + r) * os2;
    for (int d = tid; d < Dv; d += 128) {
      if (out_is_f32) ((float*)out)[oo + d * os3] = mean[d];
      else ((T*)out)[oo + d * os3] = Num<T>::from_f(mean[d]);
    }
  }
}

}  // namespace

void masked_rows_fixup(const SdpaArgs& a, const uint8_t* dead, cudaStream_t stream) {
  if ((int64_t)a.B * a.Hq * a.Lq == 0 || a.Lk == 0) return;
  OMX_CHECK(a.Dv <= 640, "[scaled_dot_product_attention] head_dim > 640 is not supported");
  dim3 grid((a.Lq + 127) / 128, a.Hq, a.B);
  const int64_t* vs = a.v->strides;
  const int64_t* os = a.out->strides;
  const int f32o = a.out->dtype == OMX_FLOAT32 ? 1 : 0;
  auto go = [&](auto* vp) {
    using T = std::remove_cv_t<std::remove_pointer_t<decltype(vp)>>;
    masked_rows_fixup_kernel<T><<<grid, 128, 0, stream>>>(dead, vp, vs[0], vs[1], vs[2], vs[3], a.out->data, os[0],
                                                          os[1], os[2], os[3], a.Hq, a.Hkv, a.Lq, a.Lk, a.Dv, f32o);
  };
  switch (a.q->dtype) {
    case OMX_FLOAT32: go((const float*)a.v->data); break;
    case OMX_BFLOAT16: go((const __nv_bfloat16*)a.v->data); break;
    default: go((const __half*)a.v->data); break;
  }
  count_launch();
  OMX_CUDA(cudaGetLastError());
}

void sdpa_generic(const SdpaArgs& a, cudaStream_t stream) {
  // up to 640: the absorbed-MLA layout of GLM-4.7-Flash (keys 512 + 64, values 512; glm-4.7-flash-mlx/src/model.rs:263-299)
  OMX_CHECK(a.D <= 640 && a.Dv <= 640, "[scaled_dot_product_attention] head_dim > 640 is not supported");
  GenParams p;
  p.q = a.q->data;
  p.k = a.k->data;
  p.v = a.v->data;
  p.out = a.out->data;
  p.mask = a.mask ? a.mask->data : nullptr;
  for (int i = 0; i < 4; ++i) {
    p.qs[i] = a.q->strides[i];
    p.ks[i] = a.k->strides[i];
    p.vs[i] = a.v->strides[i];
    p.os[i] = a.out->strides[i];
    p.ms[i] = a.mask_strides[i];
  }
  p.B = a.B; p.Hq = a.Hq; p.Hkv = a.Hkv; p.Lq = a.Lq; p.Lk = a.Lk; p.D = a.D; p.Dv = a.Dv;
  p.scale = a.scale;
  p.mask_mode = a.mask_mode;
  p.mask_is_f32 = (a.mask && a.mask->dtype == OMX_FLOAT32) ? 1 : 0;
  p.out_is_f32 = a.out->dtype == OMX_FLOAT32 ? 1 : 0;
  note_launch("sdpa_generic");
  switch (a.q->dtype) {
    case OMX_FLOAT32: launch<float>(p, stream); break;
    case OMX_BFLOAT16: launch<__nv_bfloat16>(p, stream); break;
    default: launch<__half>(p, stream); break;
  }
}

}  // namespace omx
