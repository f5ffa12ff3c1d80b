// sdpa_f32_tiled.cu -- float32 attention for Lq > 1 (prefill / DiT in float32) on the FFMA pipe.
//
// The float32 callers of mlx_rs::fast::scaled_dot_product_attention with more than one query row -- the prefill
// of the reference's own CPU-runnable configuration (Qwen3-0.6B, float32: qwen3-mlx/src/model.rs:399-408) and the
// FLUX example run in float32 (flux-klein-mlx/src/klein_model.rs:460-489) -- have to meet 1e-4 relative, which a
// TF32 tensor-core product (10-bit mantissa) does not.  sdpa_generic.cu serves them correctly with one warp per
// query row, i.e. every row re-streams K / V through L2; this kernel is the tiled spelling of the same arithmetic:
//
//   CTA = 128 threads = 64 query rows of one (batch, head); key tiles of 64; Q, one K-or-V tile and P live in
//   shared memory (85 KB at head_dim 128: two CTAs per SM, one loads while the other multiplies);
//   S = Q K^T: thread (ty, tx) owns rows ty + 16 i (i < 4) x columns tx + 8 j (j < 8): 12 LDS.128 per 128 FFMA;
//   online softmax in the log2 domain, row reductions over the 8 lanes that share a row (shuffles);
//   O += P V:  rows ty + 16 i x feature columns 4 tx + 32 j: 20 LDS.128 per 256 FFMA.
//   Pitches (D + 4 floats for Q / K / V rows, 72 for P) keep every warp-wide access conflict-free.
//
// Mask semantics are sdpa_generic's (the MLX fallback graph): causal keys past the diagonal are excluded, bool
// entries that are masked take finfo(float32).min (a fully masked row degrades to the uniform average), additive
// masks are added to the scaled scores.  Causal launches stop at the diagonal tile.
#include <algorithm>

#include "omx_common.cuh"
#include "omx_internal.h"

namespace omx {

namespace {

struct TiledParams {
  const float *q, *k, *v;
  const void* mask;
  float* out;
  int64_t qs[4], ks[4], vs[4], os[4], ms[4];
  int B, Hq, Hkv, Lq, Lk;
  float scale;
  int mask_mode;
};

constexpr int kBM = 64, kBN = 64, kPP = 72;  // rows per CTA, keys per tile, pitch of P

template <int D>
__global__ void __launch_bounds__(128, 2) sdpa_f32_tiled_kernel(const __grid_constant__ TiledParams p) {
  constexpr int DP = D + 4;   // row pitch of the Q / K / V tiles (floats)
  constexpr int ND4 = D / 32;  // float4 column groups per thread in O
  extern __shared__ float smem[];
  float* Qs = smem;                 // [64][DP]
  float* KVs = Qs + kBM * DP;       // [64][DP]  K during S = Q K^T, then V during O += P V
  float* Ps = KVs + kBN * DP;       // [64][kPP]

  const int tid = threadIdx.x, ty = tid >> 3, tx = tid & 7;
  // causal: the row blocks near the end see the most keys -- launch them first (CTAs are dispatched in index order)
  const int mb = p.mask_mode == MASK_CAUSAL ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x;
  const int m0 = mb * kBM, h = blockIdx.y, b = blockIdx.z;
  const int hk = h / (p.Hq / p.Hkv);
  const float* qg = p.q + b * p.qs[0] + h * p.qs[1];
  const float* kg = p.k + b * p.ks[0] + hk * p.ks[1];
  const float* vg = p.v + b * p.vs[0] + hk * p.vs[1];
  const int q_off = max(p.Lk - p.Lq, 0);  // bottom-right aligned causal mask (mlx: "causal")

  // rows of a [64][D] tile: 128 threads x float4, consecutive threads along a row; rows past `n_rows` are zero
  auto load_tile = [&](float* dst, const float* src, int64_t row_stride, int row0, int n_rows) {
    constexpr int PER_ROW = D / 4;
#pragma unroll 4
    for (int idx = tid; idx < 64 * PER_ROW; idx += 128) {
      const int r = idx / PER_ROW, c = idx % PER_ROW;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row0 + r < n_rows) v = __ldg(reinterpret_cast<const float4*>(src + (int64_t)(row0 + r) * row_stride) + c);
      *reinterpret_cast<float4*>(dst + r * DP + c * 4) = v;
    }
  };
  load_tile(Qs, qg, p.qs[2], m0, p.Lq);

  float o[4][ND4][4];
  float mrow[4], lrow[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    mrow[i] = -INFINITY;
    lrow[i] = 0.f;
#pragma unroll
    for (int j = 0; j < ND4; ++j) o[i][j][0] = o[i][j][1] = o[i][j][2] = o[i][j][3] = 0.f;
  }
  const int k_end = p.mask_mode == MASK_CAUSAL ? min(p.Lk, q_off + m0 + kBM) : p.Lk;  // keys this block can see
  const float fill = -3.4028234663852886e38f;  // finfo(float32).min
  const float sl2 = p.scale * kLog2e;

  for (int n0 = 0; n0 < k_end; n0 += kBN) {
    __syncthreads();  // everybody is done with the previous tile's V (and, first time, nothing)
    load_tile(KVs, kg, p.ks[2], n0, p.Lk);
    __syncthreads();
    // ---- S = Q K^T
    float s[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) s[i][j] = 0.f;
#pragma unroll 2
    for (int d = 0; d < D; d += 4) {
      float4 qv[4], kv[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) qv[i] = *reinterpret_cast<const float4*>(Qs + (ty + 16 * i) * DP + d);
#pragma unroll
      for (int j = 0; j < 8; ++j) kv[j] = *reinterpret_cast<const float4*>(KVs + (tx + 8 * j) * DP + d);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          s[i][j] = fmaf(qv[i].x, kv[j].x, s[i][j]);
          s[i][j] = fmaf(qv[i].y, kv[j].y, s[i][j]);
          s[i][j] = fmaf(qv[i].z, kv[j].z, s[i][j]);
          s[i][j] = fmaf(qv[i].w, kv[j].w, s[i][j]);
        }
    }
    // ---- scale, mask, online softmax (log2 domain); P -> shared memory
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = m0 + ty + 16 * i;
      const int64_t mrow_off = b * p.ms[0] + h * p.ms[1] + (int64_t)row * p.ms[2];
      float tmax = -INFINITY;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int col = n0 + tx + 8 * j;
        float x = s[i][j] * sl2;
        if (col >= p.Lk || row >= p.Lq || (p.mask_mode == MASK_CAUSAL && col > q_off + row)) {
          x = -INFINITY;
        } else if (p.mask_mode == MASK_BOOL) {
          if (!((const uint8_t*)p.mask)[mrow_off + col * p.ms[3]]) x = fill;
        } else if (p.mask_mode == MASK_ADD) {
          x = fmaf(((const float*)p.mask)[mrow_off + col * p.ms[3]], kLog2e, x);
        }
        s[i][j] = x;
        tmax = fmaxf(tmax, x);
      }
      tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, 1));
      tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, 2));
      tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, 4));
      const float m_new = fmaxf(mrow[i], tmax);
      const float m_safe = m_new == -INFINITY ? 0.f : m_new;
      const float corr = fast_exp2(mrow[i] - m_safe);
      float psum = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float e = fast_exp2(s[i][j] - m_safe);
        psum += e;
        Ps[(ty + 16 * i) * kPP + tx + 8 * j] = e;
      }
      psum += __shfl_xor_sync(0xffffffffu, psum, 1);
      psum += __shfl_xor_sync(0xffffffffu, psum, 2);
      psum += __shfl_xor_sync(0xffffffffu, psum, 4);
      lrow[i] = lrow[i] * corr + psum;
      mrow[i] = m_new;
#pragma unroll
      for (int j = 0; j < ND4; ++j) {
        o[i][j][0] *= corr; o[i][j][1] *= corr; o[i][j][2] *= corr; o[i][j][3] *= corr;
      }
    }
    __syncthreads();  // P complete, K no longer read
    load_tile(KVs, vg, p.vs[2], n0, p.Lk);
    __syncthreads();
    // ---- O += P V
#pragma unroll 2
    for (int k = 0; k < kBN; k += 4) {
      float4 pv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) pv[i] = *reinterpret_cast<const float4*>(Ps + (ty + 16 * i) * kPP + k);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        float4 vv[ND4];
#pragma unroll
        for (int j = 0; j < ND4; ++j) vv[j] = *reinterpret_cast<const float4*>(KVs + (k + kk) * DP + 4 * tx + 32 * j);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float pe = kk == 0 ? pv[i].x : (kk == 1 ? pv[i].y : (kk == 2 ? pv[i].z : pv[i].w));
#pragma unroll
          for (int j = 0; j < ND4; ++j) {
            o[i][j][0] = fmaf(pe, vv[j].x, o[i][j][0]);
            o[i][j][1] = fmaf(pe, vv[j].y, o[i][j][1]);
            o[i][j][2] = fmaf(pe, vv[j].z, o[i][j][2]);
            o[i][j][3] = fmaf(pe, vv[j].w, o[i][j][3]);
          }
        }
      }
    }
  }
  // ---- epilogue
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = m0 + ty + 16 * i;
    if (row >= p.Lq) continue;
    const float inv = 1.0f / lrow[i];
    float* og = p.out + b * p.os[0] + h * p.os[1] + (int64_t)row * p.os[2];
#pragma unroll
    for (int j = 0; j < ND4; ++j)
      *reinterpret_cast<float4*>(og + 4 * tx + 32 * j) =
          make_float4(o[i][j][0] * inv, o[i][j][1] * inv, o[i][j][2] * inv, o[i][j][3] * inv);
  }
}

bool rows16(const omx_array* a) {  // rows contiguous, 16-byte aligned, every outer stride a multiple of 4 floats
  if (a->strides[3] != 1 && a->shape[3] != 1) return false;
  if (reinterpret_cast<uintptr_t>(a->data) & 15) return false;
  for (int i = 0; i < 3; ++i)
    if (a->strides[i] % 4) return false;
  return true;
}

}  // namespace

bool sdpa_f32_tiled_supported(const SdpaArgs& a, const char** why) {
  auto no = [&](const char* w) {
    if (why) *why = w;
    return false;
  };
  if (a.q->dtype != OMX_FLOAT32 || a.k->dtype != OMX_FLOAT32 || a.v->dtype != OMX_FLOAT32 ||
      a.out->dtype != OMX_FLOAT32)
    return no("not float32");
  if (a.D != a.Dv || !(a.D == 64 || a.D == 128)) return no("head_dim not in {64, 128}");
  if (a.Lk < 1) return no("no keys");
  if (a.mask_mode == MASK_ADD && a.mask->dtype != OMX_FLOAT32) return no("additive mask is not float32");
  if (!rows16(a.q) || !rows16(a.k) || !rows16(a.v) || !rows16(a.out)) return no("rows not contiguous / 16-byte aligned");
  return true;
}

void sdpa_f32_tiled(const SdpaArgs& a, cudaStream_t stream) {
  TiledParams p;
  p.q = (const float*)a.q->data;
  p.k = (const float*)a.k->data;
  p.v = (const float*)a.v->data;
  p.out = (float*)a.out->data;
  p.mask = a.mask ? a.mask->data : nullptr;
  for (int i = 0; i < 4; ++i) {
    p.qs[i] = a.q->strides[i];
    p.ks[i] = a.k->strides[i];
    p.vs[i] = a.v->strides[i];
    p.os[i] = a.out->strides[i];
    p.ms[i] = a.mask_strides[i];
  }
  p.B = a.B; p.Hq = a.Hq; p.Hkv = a.Hkv; p.Lq = a.Lq; p.Lk = a.Lk;
  p.scale = a.scale;
  p.mask_mode = a.mask_mode;
  dim3 grid((a.Lq + kBM - 1) / kBM, a.Hq, a.B);
  note_launch("sdpa_f32_tiled");
  auto go = [&](auto kern, int D) {
    const size_t smem = sizeof(float) * ((size_t)(kBM + kBN) * (D + 4) + (size_t)kBM * kPP);
    OMX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, 128, smem, stream>>>(p);
  };
  if (a.D == 128) go(sdpa_f32_tiled_kernel<128>, 128);
  else go(sdpa_f32_tiled_kernel<64>, 64);
  count_launch();
  OMX_CUDA(cudaGetLastError());
}

}  // namespace omx
