// prologue.cu -- the q/k/v prologue of the attention path in ONE launch.
//
// What sits between the projections and the attention kernel in every caller is a chain of
// row-wise ops over [.., D] head rows:
//   LLM prefill (qwen3-mlx/src/model.rs:172-201): q_norm(q), k_norm(k), rope(q, off), rope(k, off),
//       cache.update_and_fetch(k', v)            -> 4 full read+write passes + 2 slice updates
//   DiT blocks (flux-klein-mlx/src/klein_model.rs:443-462,641-648; zimage-mlx/src/zimage_model.rs:345-352):
//       norm_q, norm_k (x2 streams), apply_rope (x4), concatenate [txt; img] for K and V
// All of it is per-row work, so one kernel does it: a launch carries up to 6 SEGMENTS (source view,
// destination view, optional RMSNorm weight, rotate or copy) and ONE THREAD OWNS ONE ROW -- the
// reference's rms_norm is a left-to-right f32 sum over the row, which is only reproducible bit for
// bit if the row is not split -- held in registers (D = 64 / 128); the warp moves its 32 rows between
// HBM and a shared-memory tile with coalesced 128-bit accesses.  Destinations are arbitrary strided views, so k' lands directly in the KV-cache rows,
// v in its rows, q' in scratch (prefill), or q/k/v in the joint [txt; img] buffers (DiT).
//
// Numerics = the standalone kernels' (norm.cu, rope.cu): every primitive rounds to the array dtype.
// For 16-bit types the rotation uses packed HMUL2 / HADD2 (.rn, no contraction): a product of two
// 8- or 11-bit significands is exact in f32, and sums of two such values round identically once or
// twice (24 >= 2p + 2), so T(f32(a) op f32(b)) == a op_T b bit for bit -- checked against the oracle.
// HBM-bound: read + write of every row once (C3: 1.61 GB per layer instead of 2.95 GB in 5 passes).
#include <algorithm>
#include <type_traits>

#include "omx_common.cuh"
#include "omx_internal.h"

namespace omx {

namespace {

constexpr int kMaxSeg = 6;
constexpr int kThreads = 128;
// Resident blocks per SM.  Measured on B200 (C3 prologue, gpurun_out/s6_composite.log): 2 blocks at 186
// registers 0.31 ms; 3 blocks squeezed to 168 registers (120 B of spills) 0.39 ms.
template <typename T>
constexpr int kBlocksPerSM() { return 2; }

struct Seg {
  const void* x;
  void* out;
  const void* w;         // [D] RMSNorm weight (contiguous, x's dtype) or null
  int64_t xs[3], os[3];  // element strides of (b, h, l); the feature axis is contiguous
  int H, L;
  int64_t row_end;       // exclusive prefix sum of rows (= B*H*L) over the segments
  int rope;              // 1: rotate, 0: copy (after the optional norm)
  int tok0;              // table row of l == 0: position (mode 1) or token index (mode 2)
  int64_t dyn_os;        // graph mode: destination offset in elements per position
  int paged_dst;         // paged mode: out = page pool, os[0] = page stride
  int writes_len;        // paged mode: this segment's head-0 rows publish lens_in[b] + 1
};

struct PrologueParams {
  Seg seg[kMaxSeg];
  int nseg;
  int traditional;  // mode 1 only (mode 2 is always adjacent pairs)
  int mode;         // 1: float32 tables [n_pos, half] by position (fast::rope); 2: T tables [B, S, half] (DiT)
  const float *cos, *sin;
  const void *cos16, *sin16;  // mode 1, 16-bit rows: the table rounded to the row dtype (null: round on the fly)
  int n_pos;
  const void *tcos, *tsin;
  int64_t cs[3], ss[3];
  int tvec;  // mode 2: table rows are contiguous and 16-byte aligned
  float eps, inv_n;
  const int* pos_dev;  // graph mode: device-resident position added to tok0 / the destinations
  // paged mode: per-sequence lengths (device), block table, where the new lengths go
  const int* lens_in;
  int* lens_out;
  const int* block_table;
  int bt_stride;
};

// ---- packed 16-bit arithmetic with one rounding per primitive
template <typename T>
struct P2;
template <>
struct P2<__nv_bfloat16> {
  using T2 = __nv_bfloat162;
  static __device__ __forceinline__ T2 mul(T2 a, T2 b) { return __hmul2_rn(a, b); }
  static __device__ __forceinline__ T2 add(T2 a, T2 b) { return __hadd2_rn(a, b); }
  static __device__ __forceinline__ T2 sub(T2 a, T2 b) { return __hsub2_rn(a, b); }
  static __device__ __forceinline__ T2 pack(float a, float b) { return __floats2bfloat162_rn(a, b); }
  static __device__ __forceinline__ float2 unpack(T2 a) { return __bfloat1622float2(a); }
  static __device__ __forceinline__ T2 lows(T2 a, T2 b) { return __lows2bfloat162(a, b); }
  static __device__ __forceinline__ T2 highs(T2 a, T2 b) { return __highs2bfloat162(a, b); }
  static __device__ __forceinline__ T2 make(__nv_bfloat16 a, __nv_bfloat16 b) { return __halves2bfloat162(a, b); }
};
template <>
struct P2<__half> {
  using T2 = __half2;
  static __device__ __forceinline__ T2 mul(T2 a, T2 b) { return __hmul2_rn(a, b); }
  static __device__ __forceinline__ T2 add(T2 a, T2 b) { return __hadd2_rn(a, b); }
  static __device__ __forceinline__ T2 sub(T2 a, T2 b) { return __hsub2_rn(a, b); }
  static __device__ __forceinline__ T2 pack(float a, float b) { return __floats2half2_rn(a, b); }
  static __device__ __forceinline__ float2 unpack(T2 a) { return __half22float2(a); }
  static __device__ __forceinline__ T2 lows(T2 a, T2 b) { return __lows2half2(a, b); }
  static __device__ __forceinline__ T2 highs(T2 a, T2 b) { return __highs2half2(a, b); }
  static __device__ __forceinline__ T2 make(__half a, __half b) { return __halves2half2(a, b); }
};

// ---------------------------------------------------------------- 16-bit rows
template <typename T, int D, int DIMS>
__device__ __forceinline__ void transform_row16(const PrologueParams& p, const Seg& s, uint4 (&rv)[D / 8], int b,
                                                int l) {
  using Q = P2<T>;
  using T2 = typename Q::T2;
  constexpr int NP = D / 2;  // packed registers
  T2* h2 = reinterpret_cast<T2*>(rv);
  constexpr int HALF = DIMS > 0 ? DIMS / 2 : 2;
  constexpr int HP = HALF / 2;  // packed cos / sin registers
  T2 c2[HP], s2[HP];
  const bool roped = DIMS > 0 && s.rope;
  bool early = false;
  if constexpr (DIMS > 0) {
    // the table row in the row dtype (fast::rope tables rounded once on the host, or the DiT callers' own tables),
    // requested AHEAD of the norm arithmetic: in the profile of the float32-table version 22 % of the kernel's stall
    // samples sat on the first use of these loads (32 LDG.128 per thread issued after the norm; now 16, before it)
    const T *tc = nullptr, *ts = nullptr;
    if (roped && p.mode == 1 && p.cos16 != nullptr) {
      const int pos = min(s.tok0 + l, p.n_pos - 1);
      tc = (const T*)p.cos16 + (size_t)pos * HALF;
      ts = (const T*)p.sin16 + (size_t)pos * HALF;
    } else if (roped && p.mode != 1 && p.tvec) {
      tc = (const T*)p.tcos + b * p.cs[0] + (int64_t)(s.tok0 + l) * p.cs[1];
      ts = (const T*)p.tsin + b * p.ss[0] + (int64_t)(s.tok0 + l) * p.ss[1];
    }
    early = tc != nullptr;
    if (early) {
      const uint4* cq = reinterpret_cast<const uint4*>(tc);
      const uint4* sq = reinterpret_cast<const uint4*>(ts);
#pragma unroll
      for (int i = 0; i < HALF / 8; ++i) {
        const uint4 cv = cq[i], sv = sq[i];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          c2[i * 4 + e] = reinterpret_cast<const T2*>(&cv)[e];
          s2[i * 4 + e] = reinterpret_cast<const T2*>(&sv)[e];
        }
      }
    }
  }
  if (s.w) {
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      const float2 f = Q::unpack(h2[j]);
      acc = __fadd_rn(acc, __fmul_rn(f.x, f.x));
      acc = __fadd_rn(acc, __fmul_rn(f.y, f.y));
    }
    const float rs = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(__fmul_rn(acc, p.inv_n), p.eps)));
    const uint4* wv = reinterpret_cast<const uint4*>(s.w);
#pragma unroll
    for (int i = 0; i < D / 8; ++i) {
      const uint4 wq = wv[i];
      const T2* w2 = reinterpret_cast<const T2*>(&wq);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = Q::unpack(h2[i * 4 + e]);
        h2[i * 4 + e] = Q::mul(w2[e], Q::pack(__fmul_rn(f.x, rs), __fmul_rn(f.y, rs)));
      }
    }
  }
  if constexpr (DIMS > 0) {
    if (!roped) return;
    if (early) {
      // c2 / s2 hold the row already
    } else if (p.mode == 1) {
      const int pos = min(s.tok0 + l, p.n_pos - 1);
      const float4* c4 = reinterpret_cast<const float4*>(p.cos + (size_t)pos * HALF);
      const float4* s4 = reinterpret_cast<const float4*>(p.sin + (size_t)pos * HALF);
#pragma unroll
      for (int i = 0; i < HALF / 4; ++i) {
        const float4 c = c4[i], sn = s4[i];
        c2[2 * i] = Q::pack(c.x, c.y);
        c2[2 * i + 1] = Q::pack(c.z, c.w);
        s2[2 * i] = Q::pack(sn.x, sn.y);
        s2[2 * i + 1] = Q::pack(sn.z, sn.w);
      }
    } else {
      const T* tc = (const T*)p.tcos + b * p.cs[0] + (int64_t)(s.tok0 + l) * p.cs[1];
      const T* ts = (const T*)p.tsin + b * p.ss[0] + (int64_t)(s.tok0 + l) * p.ss[1];
      // (contiguous, aligned tables took the early path above)
#pragma unroll
      for (int j = 0; j < HP; ++j) {
        c2[j] = Q::make(tc[(2 * j) * p.cs[2]], tc[(2 * j + 1) * p.cs[2]]);
        s2[j] = Q::make(ts[(2 * j) * p.ss[2]], ts[(2 * j + 1) * p.ss[2]]);
      }
    }
    if (p.mode == 1 && !p.traditional) {
      // pairs (i, i + HALF): packed register j holds x1[2j..2j+1], register HP + j holds x2[2j..2j+1]
#pragma unroll
      for (int j = 0; j < HP; ++j) {
        const T2 x1 = h2[j], x2 = h2[HP + j];
        h2[j] = Q::sub(Q::mul(x1, c2[j]), Q::mul(x2, s2[j]));
        h2[HP + j] = Q::add(Q::mul(x1, s2[j]), Q::mul(x2, c2[j]));
      }
    } else {
      // adjacent pairs: register u holds (x1, x2) of pair u; two pairs per packed operation
#pragma unroll
      for (int j = 0; j < HP; ++j) {
        const T2 r0 = h2[2 * j], r1 = h2[2 * j + 1];
        const T2 x1 = Q::lows(r0, r1), x2 = Q::highs(r0, r1);
        const T2 o1 = Q::sub(Q::mul(x1, c2[j]), Q::mul(x2, s2[j]));
        const T2 o2 = Q::add(Q::mul(x1, s2[j]), Q::mul(x2, c2[j]));
        h2[2 * j] = Q::lows(o1, o2);
        h2[2 * j + 1] = Q::highs(o1, o2);
      }
    }
  }
}

// ---------------------------------------------------------------- float32 rows
template <int D, int DIMS>
__device__ __forceinline__ void transform_row32(const PrologueParams& p, const Seg& s, uint4 (&rv)[D / 4], int b,
                                                int l) {
  float* x = reinterpret_cast<float*>(rv);
  if (s.w) {
    float acc = 0.f;
#pragma unroll
    for (int d = 0; d < D; ++d) acc = __fadd_rn(acc, __fmul_rn(x[d], x[d]));
    const float rs = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(__fmul_rn(acc, p.inv_n), p.eps)));
    const float4* wv = reinterpret_cast<const float4*>(s.w);
#pragma unroll
    for (int i = 0; i < D / 4; ++i) {
      const float4 w = wv[i];
      x[4 * i] = __fmul_rn(w.x, __fmul_rn(x[4 * i], rs));
      x[4 * i + 1] = __fmul_rn(w.y, __fmul_rn(x[4 * i + 1], rs));
      x[4 * i + 2] = __fmul_rn(w.z, __fmul_rn(x[4 * i + 2], rs));
      x[4 * i + 3] = __fmul_rn(w.w, __fmul_rn(x[4 * i + 3], rs));
    }
  }
  if constexpr (DIMS > 0) {
    if (!s.rope) return;
    constexpr int HALF = DIMS / 2;
    const float *cr, *sr;
    int64_t cst = 1, sst = 1;
    if (p.mode == 1) {
      const int pos = min(s.tok0 + l, p.n_pos - 1);
      cr = p.cos + (size_t)pos * HALF;
      sr = p.sin + (size_t)pos * HALF;
    } else {
      cr = (const float*)p.tcos + b * p.cs[0] + (int64_t)(s.tok0 + l) * p.cs[1];
      sr = (const float*)p.tsin + b * p.ss[0] + (int64_t)(s.tok0 + l) * p.ss[1];
      cst = p.cs[2];
      sst = p.ss[2];
    }
    if (p.mode == 1 && !p.traditional) {
#pragma unroll
      for (int u = 0; u < HALF; ++u)
        rope_pair<float>(x[u], x[u + HALF], cr[u], sr[u], x[u], x[u + HALF]);
    } else {
#pragma unroll
      for (int u = 0; u < HALF; ++u)
        rope_pair<float>(x[2 * u], x[2 * u + 1], cr[u * cst], sr[u * sst], x[2 * u], x[2 * u + 1]);
    }
  }
}

// Global accesses are made by the WARP, not by the row owner: a warp's 32 rows arrive through
// fully coalesced 16-byte cp.async copies (NV lanes per row, 32 / NV rows per instruction) in a
// padded shared-memory tile, each thread then takes ITS row from the tile (conflict-free: the
// 16-byte pitch offset walks the banks), transforms it in registers, puts it back, and the warp
// stores the tile with the same coalesced pattern to the per-row destinations.  Warps are
// persistent and double-buffered: tile n+1 is in flight (no registers held) while tile n is being
// transformed, so a few warps per SM keep HBM busy.  Warp-local: no block barrier anywhere.
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}

template <typename T, int D, int DIMS>
__global__ void __launch_bounds__(kThreads, kBlocksPerSM<T>()) qkv_prologue_kernel(const __grid_constant__ PrologueParams p,
                                                                const int64_t n_tiles) {
  constexpr int NV = D * (int)sizeof(T) / 16;  // 16-byte chunks per row (8, 16 or 32)
  constexpr int RPI = 32 / NV;                 // rows per warp-wide access
  constexpr int PITCH = NV + 1;
  extern __shared__ uint4 tile_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint4* tiles = tile_raw + (size_t)warp * 2 * 32 * PITCH;
  const int sub = lane / NV, chunk = lane % NV;
  const int64_t stride = (int64_t)gridDim.x * (kThreads / 32);
  const int dpos = p.pos_dev ? *p.pos_dev : 0;  // graph mode: the step's position lives in device memory

  struct RowRef {
    int si, b, l;  // si < 0: no row
    unsigned long long dst;
  };
  // locate this lane's row of tile t, start the tile's copies into buffer `buf`
  auto fetch = [&](int64_t t, int buf) -> RowRef {
    RowRef rf{-1, 0, 0, 0ull};
    unsigned long long src = 0ull;
    if (t < n_tiles) {
      const int64_t r = t * 32 + lane;
      int si = 0;
      while (si < p.nseg && r >= p.seg[si].row_end) ++si;
      if (si < p.nseg) {
        const Seg& s = p.seg[si];
        // h fastest: the 32 rows of a warp are adjacent heads of one token in the callers'
        // [B, L, H, D] storage (contiguous) and share one cos / sin table row
        int64_t rr = r - (si ? p.seg[si - 1].row_end : 0);
        const int h = (int)(rr % s.H);
        rr /= s.H;
        rf.si = si;
        rf.l = (int)(rr % s.L);
        rf.b = (int)(rr / s.L);
        src = (unsigned long long)((const T*)s.x + rf.b * s.xs[0] + h * s.xs[1] + rf.l * s.xs[2]);
        rf.dst = (unsigned long long)((T*)s.out + rf.b * s.os[0] + h * s.os[1] + rf.l * s.os[2] + dpos * s.dyn_os);
        if (p.lens_in) {  // paged single-token step: the sequence's own length is its position
          const int len = p.lens_in[rf.b];
          if (len < 0) {  // released slot: no row
            rf.si = -1;
            src = 0ull;
            rf.dst = 0ull;
          } else {
            if (s.paged_dst) {
              const int page = p.block_table[(int64_t)rf.b * p.bt_stride + (len >> 6)];
              rf.dst = (unsigned long long)((T*)s.out + (int64_t)page * s.os[0] + h * s.os[1] + (int64_t)(len & 63) * s.os[2]);
            }
            if (s.writes_len && h == 0) p.lens_out[rf.b] = len + 1;
            rf.l += len;  // (the destination is settled: from here on l is the rope row)
          }
        }
      }
    }
    uint4* tile = tiles + buf * 32 * PITCH;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int row = i * RPI + sub;
      const unsigned long long ptr = __shfl_sync(0xffffffffu, src, row);
      if (ptr) cp_async16(&tile[row * PITCH + chunk], reinterpret_cast<const uint4*>(ptr) + chunk);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    return rf;
  };

  int64_t t = (int64_t)blockIdx.x * (kThreads / 32) + warp;
  int buf = 0;
  RowRef cur = fetch(t, 0);
  for (; t < n_tiles; t += stride, buf ^= 1) {
    const RowRef nxt = fetch(t + stride, buf ^ 1);
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncwarp();
    uint4* tile = tiles + buf * 32 * PITCH;
    if (cur.si >= 0) {
      const Seg& s = p.seg[cur.si];
      uint4 rv[NV];
#pragma unroll
      for (int i = 0; i < NV; ++i) rv[i] = tile[lane * PITCH + i];
      if constexpr (std::is_same<T, float>::value) transform_row32<D, DIMS>(p, s, rv, cur.b, cur.l + dpos);
      else transform_row16<T, D, DIMS>(p, s, rv, cur.b, cur.l + dpos);
#pragma unroll
      for (int i = 0; i < NV; ++i) tile[lane * PITCH + i] = rv[i];
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int row = i * RPI + sub;
      const unsigned long long ptr = __shfl_sync(0xffffffffu, cur.dst, row);
      if (ptr) reinterpret_cast<uint4*>(ptr)[chunk] = tile[row * PITCH + chunk];
    }
    __syncwarp();  // the tile is free again before the next iteration's copies land in it
    cur = nxt;
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

template <typename T>
bool launch_t(const PrologueParams& p, int D, int dims, int64_t rows, cudaStream_t stream) {
  const int64_t n_tiles = (rows + 31) / 32;
  const size_t smem = 2 * (size_t)kThreads * ((size_t)D * sizeof(T) / 16 + 1) * 16;
  // persistent warps: as many blocks as stay resident, never more than the work
  const int resident = smem > 112 * 1024 ? 1 : kBlocksPerSM<T>();  // (512-byte rows: one block's tiles fill the SM)
  const unsigned blocks = (unsigned)std::min<int64_t>((n_tiles + kThreads / 32 - 1) / (kThreads / 32),
                                                      (int64_t)sm_count() * resident);
#define OMX_PRO(DD, RR)                                                                                  \
  if (D == DD && dims == RR) {                                                                           \
    auto kern = qkv_prologue_kernel<T, DD, RR>;                                                          \
    if (smem > 48 * 1024)                                                                                \
      OMX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));      \
    kern<<<blocks, kThreads, smem, stream>>>(p, n_tiles);                                                \
    return true;                                                                                         \
  }
  OMX_PRO(128, 128)
  OMX_PRO(128, 64)
  OMX_PRO(128, 0)
  OMX_PRO(64, 64)
  OMX_PRO(64, 0)
  if constexpr (sizeof(T) == 2) {  // 512-byte rows (Qwen3.5: head dim 256, rope on the first 64 features)
    OMX_PRO(256, 64)
    OMX_PRO(256, 0)
  }
#undef OMX_PRO
  return false;
}

}  // namespace

bool qkv_prologue(const PrologueCall& c, cudaStream_t stream) {
  if (c.nseg < 1 || c.nseg > kMaxSeg) return false;
  const int dt = c.seg[0].x->dtype;
  const int D = (int)c.seg[0].x->shape[3];
  if (!(D == 128 || D == 64 || (D == 256 && dtype_size(dt) == 2))) return false;
  bool any_rope = false;
  for (int i = 0; i < c.nseg; ++i) any_rope = any_rope || c.seg[i].rope;
  const int dims = any_rope ? c.dims : 0;
  if (D == 256 ? !(dims == 0 || dims == 64) : !(dims == 0 || dims == D || (D == 128 && dims == 64))) return false;
  if (c.mode == 2 && any_rope && dims != D) return false;
  const int64_t ve = (int64_t)(16 / dtype_size(dt));
  PrologueParams p{};
  int64_t rows = 0;
  for (int i = 0; i < c.nseg; ++i) {
    const PrologueSeg& a = c.seg[i];
    const omx_array* x = a.x;
    const omx_array* o = &a.out;
    if (x->ndim != 4 || o->ndim != 4 || x->dtype != dt || o->dtype != dt) return false;
    if (x->shape[3] != D || o->shape[3] != D) return false;
    for (int j = 0; j < 3; ++j)
      if (x->shape[j] != o->shape[j]) return false;
    const bool inner_ok = (x->strides[3] == 1 || D == 1) && (o->strides[3] == 1 || D == 1);
    if (!inner_ok || !aligned16(x->data) || !aligned16(o->data)) return false;
    for (int j = 0; j < 3; ++j)
      if (x->shape[j] > 1 && (x->strides[j] % ve || o->strides[j] % ve)) return false;
    Seg& s = p.seg[i];
    s.x = x->data;
    s.out = o->data;
    s.w = nullptr;
    if (a.w && a.w->data) {
      if (a.w->ndim != 1 || a.w->shape[0] != D || a.w->dtype != dt || a.w->strides[0] != 1 || !aligned16(a.w->data))
        return false;
      s.w = a.w->data;
    }
    for (int j = 0; j < 3; ++j) {
      s.xs[j] = x->strides[j];
      s.os[j] = o->strides[j];
    }
    s.H = (int)x->shape[1];
    s.L = (int)x->shape[2];
    rows += x->shape[0] * x->shape[1] * x->shape[2];
    s.row_end = rows;
    s.rope = a.rope ? 1 : 0;
    s.tok0 = a.tok0;
    s.dyn_os = a.dyn_row_stride;
    s.paged_dst = a.paged_dst ? 1 : 0;
    s.writes_len = 0;
  }
  p.nseg = c.nseg;
  p.traditional = c.traditional ? 1 : 0;
  p.mode = c.mode;
  p.eps = c.eps;
  p.inv_n = 1.0f / (float)D;
  p.pos_dev = c.pos_dev;
  if (c.pos_dev && c.mode != 1) return false;
  if (c.paged) {
    if (c.mode != 1 || c.pos_dev) return false;
    p.lens_in = c.paged->lens_in;
    p.lens_out = c.paged->lens_out;
    p.block_table = c.paged->block_table;
    p.bt_stride = c.paged->bt_stride;
    bool marked = false;  // the first paged segment (the keys) publishes the new lengths
    for (int i = 0; i < c.nseg; ++i) {
      if (c.seg[i].x->shape[2] != 1) return false;  // single-token steps only
      if (c.seg[i].paged_dst && !marked) {
        p.seg[i].writes_len = 1;
        marked = true;
      }
    }
    if (!marked) return false;
  }
  if (any_rope && c.mode == 1) {
    if (!c.table.cos || c.table.half != dims / 2) return false;
    p.cos = c.table.cos;
    p.sin = c.table.sin;
    p.n_pos = c.table.n_pos;
    if (dt == OMX_BFLOAT16 || dt == OMX_FLOAT16) {
      p.cos16 = c.table.cos16[dt == OMX_BFLOAT16 ? 0 : 1];
      p.sin16 = c.table.sin16[dt == OMX_BFLOAT16 ? 0 : 1];
      if (!p.cos16 || !p.sin16) p.cos16 = p.sin16 = nullptr;
    }
  } else if (any_rope) {
    if (!c.tcos || !c.tsin || c.tcos->dtype != dt || c.tsin->dtype != dt) return false;
    p.tcos = c.tcos->data;
    p.tsin = c.tsin->data;
    for (int j = 0; j < 3; ++j) {
      p.cs[j] = c.tcs[j];
      p.ss[j] = c.tss[j];
    }
    p.tvec = c.tcs[2] == 1 && c.tss[2] == 1 && aligned16(p.tcos) && aligned16(p.tsin) && c.tcs[0] % ve == 0 &&
             c.tcs[1] % ve == 0 && c.tss[0] % ve == 0 && c.tss[1] % ve == 0 && (D / 2) % ve == 0;
  }
  if (rows == 0) return true;
  bool ok;
  switch (dt) {
    case OMX_FLOAT32: ok = launch_t<float>(p, D, dims, rows, stream); break;
    case OMX_BFLOAT16: ok = launch_t<__nv_bfloat16>(p, D, dims, rows, stream); break;
    case OMX_FLOAT16: ok = launch_t<__half>(p, D, dims, rows, stream); break;
    default: return false;
  }
  if (!ok) return false;
  count_launch();
  OMX_CUDA(cudaGetLastError());
  return true;
}

}  // namespace omx
