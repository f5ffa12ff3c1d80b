// sdpa_mma.cu -- 16-bit attention for the head shapes the tcgen05 / TMA kernels do not take, on tensor cores.
//
// Serves omx_fast_scaled_dot_product_attention (replaces mlx_fast_scaled_dot_product_attention,
// mlx-c/mlx/c/fast.h:189-198) for bf16 / f16 calls with
//   * keys wider than values -- the absorbed MLA of GLM-4.7-Flash: queries [B,20,L,576], keys [B,1,S,512+64],
//     values [B,1,S,512], ONE shared kv head (glm-4.7-flash-mlx/src/model.rs:263-299), decode and prefill;
//   * head dims outside {64, 128}: 72 / 80 (vision towers), 256 (qwen3.5), 16 / 32 (the reference's test shapes);
//   * single-token (decode) calls with grouped query heads at those head dims (Qwen3.5: 16 / 2 heads, head dim 256,
//     qwen3.5-35B-mlx/src/attention.rs), which the CUDA-core split-K kernel served at 1 TB/s;
//   * layouts / masks the specialised kernels refuse.
// r01 / early r02 sent the multi-row ones to sdpa_generic (one warp per query row on CUDA cores).
//
// Shape of the kernel: `mma.sync.m16n8k16` (f32 accumulate) fed by `ldmatrix` from shared memory, `cp.async`
// double-buffered K / V tiles.  A CTA owns 64 packed query rows of one (batch, kv head): the rows of ALL query
// heads of the kv head's group are packed token-major (row = token * G + head-in-group), so a key tile is read
// once for the whole group -- for the MLA layout that is all 20 heads.  Four row groups of 16 rows; value widths
// above 256 put a second warp on each row group: each owns half of the output columns and multiplies half of the
// FEATURES of QK^T, the partial score tiles cross through shared memory at a named barrier (both add them in the
// same order: identical softmax states).  With few packed rows (decode) row groups turn into KEY groups (KS = 2 /
// 4): each warp scores its share of a tile's keys with its own running (m, l, O), merged once at the end.  Few CTAs
// -> the key range is split over CTAs and a second launch merges the partials; in graph mode the key count and
// the split plan come from a device-resident position.  Mask semantics are sdpa_generic's (= the MLX fallback
// graph): masked bool entries take finfo(T).min, additive masks are added to the scaled scores (each rounded to
// the array dtype like the reference's op chain), causal is bottom-right aligned.
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "omx_common.cuh"
#include "omx_internal.h"

namespace omx {

namespace {

constexpr int kBM = 64;  // packed query rows per CTA

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int n = valid ? 16 : 0;  // 0 source bytes: the 16 destination bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

struct MmaParams {
  const void *q, *k, *v, *mask;
  void* out;
  float* part;  // split partials: per (b, kv head, row tile, split): [64][DVP] O, [64] m, [64] l
  int64_t qs[3], ks[3], vs[3], os[3], ms[4];
  int B, Hq, Hkv, G, Lq, Lk, D, Dv, R, MT, nsplit;
  float scale;
  int mask_mode, mask_is_f32, out_is_f32;
  // graph mode (single-token steps replayed from a CUDA graph): the key count is *pos_dev + 1, read by the kernel; Lk
  // above is the number of pinned rows, nsplit the grid's split count (sized for all of them); the kernel derives the
  // split count an eager call at that key count would use, so both produce the same bits
  const int* pos_dev;
  int split_cap, min_tiles;
  size_t part_bytes;  // graph mode: size of `part`
  // paged mode (single-token steps on the paged cache): k / v are the page pools ([page][Hkv][64][D]: ks[0] / vs[0] =
  // page stride), sequence b attends lens_in[b] + 1 keys (< 0: released slot), key row j lives in page
  // block_table[b * bt_stride + j / 64] at row j % 64
  const int* lens_in;
  const int* block_table;
  int bt_stride;
};

// splits of the key range a launch uses: as many as there are free CTA slots, at least min_tiles tiles each
__host__ __device__ inline int mma_plan_splits(int lk, int bn, int split_cap, int min_tiles) {
  const int nt = (lk + bn - 1) / bn;
  return max(1, min(split_cap, nt / min_tiles));
}

__device__ __forceinline__ void ldsm4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldsm4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}

template <typename T>
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1);
template <>
__device__ __forceinline__ void mma16816<__nv_bfloat16>(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <>
__device__ __forceinline__ void mma16816<__half>(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <typename T>
__device__ __forceinline__ uint32_t pack2(float lo, float hi);
template <>
__device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
template <>
__device__ __forceinline__ uint32_t pack2<__half>(float lo, float hi) {
  __half2 v = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

template <int DKP, int DVP, int KS>
struct MmaCfg {
  // KS key groups: with few packed rows (decode) the four warps of a column group are not four row groups of 16 rows
  // but RG = 4 / KS row groups x KS key groups -- each warp scores BN / KS keys of a tile with its own running (m, l, O),
  // merged once at the end (KS = 2: <= 32 rows, KS = 4: <= 16 rows; not for the 576 / 512 widths, whose 32-key tiles
  // leave 8 keys per warp).
  static constexpr int RG = 4 / KS;
  // warps per row group, each owning DVP / NC output columns: wide values need two (a 16 x 256 float32 fragment is
  // 128 registers); KS = 4 at 256 columns takes two as well (8 warps per CTA: its stages leave room for one CTA per
  // SM only).  Measured alternative for the 576 / 512 decode variant: four column groups (16 warps of 128 registers,
  // features split four ways) -- slower, B64 ctx 4096 198 us against 171 us (profiles/r02_mma.md).
  static constexpr int NC = DVP > 256 ? 2 : (KS == 4 && DVP == 256 ? 2 : 1);
  static constexpr int NT = 128 * NC;              // threads
  static constexpr int MINB = DKP <= 96 ? 3 : (DKP <= 128 ? 2 : 1);  // resident CTAs per SM the register budget aims at
#ifndef OMX_MMA_KS4_BN_NARROW
#define OMX_MMA_KS4_BN_NARROW 128
#endif
  // keys per tile (>= 16 per key group); narrow heads with four key groups take longer tiles: a 64-key tile of 64-wide
  // rows is little work per warp between two CTA barriers
  // (B32 ctx 4096, four key groups: 64 wide 35.2 -> 33.8 us, one head per kv head 146 -> 124 us; 32 wide 29.2 -> 25.6;
  // 80 wide 64 -> 57; 96 wide LOSES, 47.7 -> 63 us -- its stages would halve the resident CTAs -- and keeps 64 keys)
  static constexpr int BN = DKP >= 256 ? (KS == 4 ? 64 : 32) : (KS == 4 && DKP <= 80 ? OMX_MMA_KS4_BN_NARROW : 64);
  static constexpr int KP = DKP + 8;               // row pitches in elements: +16 bytes keeps ldmatrix conflict-free
  static constexpr int VP = DVP + 8;
  static constexpr int WN = DVP / NC;
  static constexpr int QR = 16 * RG;               // query rows held in shared memory
  // K / V stages: the key-group variants of the 128- and 256-wide configurations keep two tiles in flight behind the
  // one being multiplied (ncu on the Qwen3.5 decode shape with two stages: 1.45 long-scoreboard stalls per issue at
  // 22 % issue activity -- the warps waited for the next tile); the other variants have no room or no need
#ifndef OMX_MMA_NS3
#define OMX_MMA_NS3 1
#endif
  static constexpr int NS = (OMX_MMA_NS3 && KS > 1 && (DKP == 256 || DKP == 128)) ? 3 : 2;
  static_assert(BN / KS >= 16 && (KS == 1 || KS == 2 || KS == 4), "key groups");
  // several warps on a row group split the FEATURES of QK^T and exchange partial score tiles (one slot per warp:
  // BN / KS / 2 words per lane).  576 / 512 prefill: 216,064 + 16,384 = 232,448 bytes, the whole opt-in maximum.
  static constexpr bool kSplitD = NC > 1;
  static constexpr size_t smem = sizeof(uint16_t) * ((size_t)QR * KP + NS * (size_t)BN * KP + NS * (size_t)BN * VP) +
                                 (kSplitD ? (NT / 32) * (BN / KS / 2) * 32 * sizeof(float) : 0);
  static_assert(smem <= 232448, "shared memory per CTA");
};

template <typename T, int DKP, int DVP, int KS>
__global__ void __launch_bounds__(MmaCfg<DKP, DVP, KS>::NT, MmaCfg<DKP, DVP, KS>::MINB) sdpa_mma_kernel(const __grid_constant__ MmaParams p) {
  using C = MmaCfg<DKP, DVP, KS>;
  constexpr int BN = C::BN, KP = C::KP, VP = C::VP, WN = C::WN, NT = C::NT;
  constexpr int BNW = BN / KS;  // keys of a tile one warp scores
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int kQR = C::QR;
  // several warps on a row group (wide values): each multiplies its SHARE of the features and the partial score
  // tiles are exchanged through shared memory, instead of all repeating the whole QK^T product
  constexpr bool kSplitD = C::kSplitD;
  constexpr int NC = C::NC;
  T* Qs = reinterpret_cast<T*>(smem_raw);  // [kQR][KP]
  T* Ks = Qs + kQR * KP;                   // [NS][BN][KP]
  constexpr int NS = C::NS;
  T* Vs = Ks + NS * BN * KP;               // [NS][BN][VP]
  float* xbuf = reinterpret_cast<float*>(Vs + NS * BN * VP);  // kSplitD: [warps][BNW / 2 words][32 lanes]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int RG = C::RG;
  const int wr = warp & (RG - 1), wk = (warp & 3) / RG, wc = warp >> 2;
  const int g = lane >> 2, t = lane & 3;
  const int mt = blockIdx.x / p.nsplit, split = blockIdx.x - mt * p.nsplit, hk = blockIdx.y, b = blockIdx.z;
  const int m0 = mt * kBM;
  const bool dyn = p.pos_dev != nullptr || p.lens_in != nullptr;
  const int Lk = p.lens_in ? min(p.lens_in[b] + 1, p.Lk) : (p.pos_dev ? min(*p.pos_dev + 1, p.Lk) : p.Lk);  // (<= 0: released slot)
  const int nsplit = dyn ? (Lk > 0 ? mma_plan_splits(Lk, BN, p.split_cap, p.min_tiles) : 0) : p.nsplit;
  if (split >= nsplit) {  // (graph mode, short context: this CTA's split does not exist; mark its partial empty)
    float* part = p.part + ((((int64_t)b * p.Hkv + hk) * p.MT + mt) * p.nsplit + split) * (kBM * DVP + 2 * kBM);
    if (tid < kBM) {
      part[kBM * DVP + tid] = -INFINITY;
      part[kBM * DVP + kBM + tid] = 0.f;
    }
    return;
  }
  const int q_off = max(Lk - p.Lq, 0);

  // keys this CTA visits: causal launches stop at the diagonal of the tile's last token
  int lk_eff = Lk;
  if (p.mask_mode == MASK_CAUSAL) {
    const int r_last = min(m0 + kBM, p.R) - 1;
    lk_eff = min(Lk, q_off + r_last / p.G + 1);
  }
  const int nt_all = (lk_eff + BN - 1) / BN;
  const int t0 = (int)((int64_t)nt_all * split / nsplit), t1 = (int)((int64_t)nt_all * (split + 1) / nsplit);

  const T* qg = (const T*)p.q + b * p.qs[0];
  const bool paged = p.block_table != nullptr;
  const int* bt = paged ? p.block_table + (int64_t)b * p.bt_stride : nullptr;
  const T* kg = (const T*)p.k + (paged ? 0 : b * p.ks[0]) + hk * p.ks[1];
  const T* vg = (const T*)p.v + (paged ? 0 : b * p.vs[0]) + hk * p.vs[1];

  // ---- Q tile (rows beyond R and features beyond D are zero-filled)
  // (Measured alternative: one cp.async.bulk per row with mbarrier completion -- 20 % slower at 1152-byte rows, 70 %
  // slower at 160-byte rows: the per-request cost of small bulk copies; a tensor-map box per tile would need
  // swizzled, unpadded stages.)
  constexpr int CH = DKP / 8;
  for (int c = tid; c < kQR * CH; c += NT) {
    const int row = c / CH, ch = c - row * CH;
    const int r = m0 + row;
    const bool ok = r < p.R && ch * 8 < p.D;
    const int tok = ok ? r / p.G : 0, hg = ok ? r - tok * p.G : 0;
    const T* src = ok ? qg + (int64_t)(hk * p.G + hg) * p.qs[1] + (int64_t)tok * p.qs[2] + ch * 8 : qg;
    cp_async16(smem_u32(Qs + row * KP + ch * 8), src, ok);
  }
  cp_async_commit();
  // `full`: every row and every chunk of the tile exists (all but a ragged last tile, widths == the configuration's)
  // -- no predicates, no selects, immediate offsets from one running pointer per row.
  auto load_rows = [&](T* dst, const T* src, int64_t row_stride, int64_t page_stride, int j0, int nfeat, bool full,
                       auto pitch_c, auto chunks_c) {
    constexpr int PITCH = decltype(pitch_c)::value, CHK = decltype(chunks_c)::value;
    if constexpr (CHK >= 32) {  // a warp per key row, lanes over the row's chunks
      constexpr int NW = NT / 32;
      const T* rp = src + (int64_t)(j0 + warp) * row_stride + lane * 8;
      const uint32_t d0 = smem_u32(dst + warp * PITCH + lane * 8);
      // (pinning the stride in a register -- its constant-bank load was the hottest stall site in ncu -- measured neutral)
      const int64_t row_step = (int64_t)NW * row_stride;
      if (full) {
#pragma unroll
        for (int i = 0; i < BN / NW; ++i, rp += row_step) {
#pragma unroll
          for (int cb = 0; cb < CHK; cb += 32)
            if (cb + 32 <= CHK || lane < CHK - cb) cp_async16(d0 + (i * NW * PITCH + cb * 8) * 2, rp + cb * 8, true);
        }
      } else {
        for (int i = 0; i < BN / NW; ++i, rp += row_step) {
          const int j = j0 + warp + i * NW;
          const bool rok_ = j < Lk;
          const T* rq = rp;
          if (paged) rq = rok_ ? src + (int64_t)bt[j >> 6] * page_stride + (int64_t)(j & 63) * row_stride + lane * 8 : src;
#pragma unroll
          for (int cb = 0; cb < CHK; cb += 32) {
            const bool ok = rok_ && (cb + lane) * 8 < nfeat;
            if (cb + lane < CHK) cp_async16(d0 + (i * NW * PITCH + cb * 8) * 2, ok ? rq + cb * 8 : src, ok);
          }
        }
      }
    } else {
      if (full) {
#pragma unroll
        for (int c0 = 0; c0 < BN * CHK; c0 += NT) {
          const int c = c0 + tid;
          const int row = c / CHK, ch = c - row * CHK;
          if (c0 + NT <= BN * CHK || c < BN * CHK)
            cp_async16(smem_u32(dst + row * PITCH + ch * 8), src + (int64_t)(j0 + row) * row_stride + ch * 8, true);
        }
      } else {
        for (int c = tid; c < BN * CHK; c += NT) {
          const int row = c / CHK, ch = c - row * CHK;
          const int j = j0 + row;
          const bool ok = j < Lk && ch * 8 < nfeat;
          const T* rq = src;
          if (ok) rq = paged ? src + (int64_t)bt[j >> 6] * page_stride + (int64_t)(j & 63) * row_stride + ch * 8
                             : src + (int64_t)j * row_stride + ch * 8;
          cp_async16(smem_u32(dst + row * PITCH + ch * 8), rq, ok);
        }
      }
    }
  };
  const bool own_width = p.D == DKP && p.Dv == DVP;
  auto load_kv = [&](int tile, int stage) {
    const bool full = own_width && (tile + 1) * BN <= Lk && !paged;
    load_rows(Ks + stage * BN * KP, kg, p.ks[2], p.ks[0], tile * BN, p.D, full, std::integral_constant<int, KP>{},
              std::integral_constant<int, DKP / 8>{});
    load_rows(Vs + stage * BN * VP, vg, p.vs[2], p.vs[0], tile * BN, p.Dv, full, std::integral_constant<int, VP>{},
              std::integral_constant<int, DVP / 8>{});
  };
  // one commit group per tile slot, empty when the tile does not exist: the group count is the same on every path
#pragma unroll
  for (int i = 0; i < NS - 1; ++i) {
    if (t0 + i < t1) load_kv(t0 + i, i);
    cp_async_commit();
  }
  // The reference scales the queries first -- T(T(scale) * q) -- and multiplies those by the keys (mlx fast.cpp
  // fallback graph; oracle/omx_oracle.c:257-275).  Every thread rescales the chunks it fetched itself.
  cp_async_wait<NS - 1>();
  {
    const float sc = rnd<T>(p.scale);
    for (int c = tid; c < kQR * CH; c += NT) {
      const int row = c / CH, ch = c - row * CH;
      if (m0 + row >= p.R || ch * 8 >= p.D) continue;
      uint4* qp = reinterpret_cast<uint4*>(Qs + row * KP + ch * 8);
      uint4 v = *qp;
      T* e = reinterpret_cast<T*>(&v);
#pragma unroll
      for (int i = 0; i < 8; ++i) e[i] = Num<T>::from_f(__fmul_rn(sc, Num<T>::to_f(e[i])));
      *qp = v;
    }
  }

  // ---- the two rows this thread holds in every accumulator fragment
  int rtok[2], rjmax[2];
  int64_t mrow[2];
  bool rok[2];
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int r = m0 + wr * 16 + g + 8 * e;
    rok[e] = r < p.R;
    const int tok = rok[e] ? r / p.G : 0, hg = rok[e] ? r - tok * p.G : 0;
    rtok[e] = tok;
    rjmax[e] = p.mask_mode == MASK_CAUSAL ? min(Lk, q_off + tok + 1) : Lk;
    mrow[e] = b * p.ms[0] + (int64_t)(hk * p.G + hg) * p.ms[1] + (int64_t)tok * p.ms[2];
  }
  const float fill = Num<T>::lowest();
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  float o[WN / 8][4];
#pragma unroll
  for (int n = 0; n < WN / 8; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;

  const bool warp_rows = m0 + wr * 16 < p.R;
  const int nk = (p.D + 15) >> 4;  // k16 steps that hold real features
  // per-lane ldmatrix offsets (elements)
  const int a_off = (wr * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * KP + (lane >> 4) * 8;
  const int bk_off = (wk * BNW + (lane & 7) + (lane >> 4) * 8) * KP + ((lane >> 3) & 1) * 8;
  const int bv_off = (wk * BNW + (lane & 7) + ((lane >> 3) & 1) * 8) * VP + wc * WN + (lane >> 4) * 8;
  const uint32_t q_base = smem_u32(Qs) + 2 * a_off;

  for (int tile = t0; tile < t1; ++tile) {
    const int it = tile - t0, stage = it % NS;
    // the stage freed by the barrier that closed the previous step takes tile + NS - 1
    if (tile + NS - 1 < t1) load_kv(tile + NS - 1, (it + NS - 1) % NS);
    cp_async_commit();
    cp_async_wait<NS - 1>();  // all but the NS - 1 youngest groups: this tile has landed
    __syncthreads();
    if (warp_rows) {  // (a row group past the last packed row only helps with the loads)
    const uint32_t k_base = smem_u32(Ks + stage * BN * KP) + 2 * bk_off;
    const uint32_t v_base = smem_u32(Vs + stage * BN * VP) + 2 * bv_off;

    // ---- S = Q K^T for this warp's 16 rows x BNW keys
    float s[BNW / 8][4];
#pragma unroll
    for (int n = 0; n < BNW / 8; ++n) s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
    {
      // fragments of step kk + 1 are requested before the MMAs of step kk issue (the asm statements keep their order)
      constexpr int NPQ = BNW / 16;
      uint32_t a0[4], a1[4], b0[NPQ][4], b1[NPQ][4];
      // k16 steps [kb, kb + kc) are this warp's (all of them unless the features are split over the warp pair)
      constexpr int CNT = DKP / 16 / (kSplitD ? NC : 1);
      static_assert(DKP / 16 % (kSplitD ? NC : 1) == 0, "feature shares");
      int kb = 0, kc = nk;
      if constexpr (kSplitD) {
        const int share = nk == DKP / 16 ? CNT : (nk + NC - 1) / NC;
        kb = wc * share;
        kc = max(0, min(nk, kb + share) - kb);
      }
      const uint32_t q_b = q_base + kb * 32, k_b = k_base + kb * 32;
      auto fetch = [&](int kk, uint32_t (&a)[4], uint32_t (&bb)[NPQ][4]) {
        ldsm4(q_b + kk * 32, a);
#pragma unroll
        for (int np = 0; np < NPQ; ++np) ldsm4(k_b + (np * 16 * KP) * 2 + kk * 32, bb[np]);
      };
      auto mult = [&](const uint32_t (&a)[4], const uint32_t (&bb)[NPQ][4]) {
#pragma unroll
        for (int np = 0; np < NPQ; ++np) {
          mma16816<T>(s[2 * np], a, bb[np][0], bb[np][1]);
          mma16816<T>(s[2 * np + 1], a, bb[np][2], bb[np][3]);
        }
      };
      if (nk == DKP / 16) {
        // the configuration's own width (MLA: 36 steps): straight-line code, immediate offsets; with few key
        // columns per warp the even / odd steps feed separate accumulators (chains half as long)
        constexpr bool kTwo = BNW / 8 <= 2 && CNT >= 12;
        float s2[kTwo ? BNW / 8 : 1][4];
#pragma unroll
        for (int n = 0; n < (kTwo ? BNW / 8 : 1); ++n) s2[n][0] = s2[n][1] = s2[n][2] = s2[n][3] = 0.f;
        fetch(0, a0, b0);
#pragma unroll
        for (int kk = 0; kk < CNT; kk += 2) {
          if (kk + 1 < CNT) fetch(kk + 1, a1, b1);
          mult(a0, b0);
          if (kk + 1 < CNT) {
            if (kk + 2 < CNT) fetch(kk + 2, a0, b0);
            if constexpr (kTwo) {
#pragma unroll
              for (int np = 0; np < NPQ; ++np) {
                mma16816<T>(s2[2 * np], a1, b1[np][0], b1[np][1]);
                mma16816<T>(s2[2 * np + 1], a1, b1[np][2], b1[np][3]);
              }
            } else {
              mult(a1, b1);
            }
          }
        }
        if constexpr (kTwo) {
#pragma unroll
          for (int n = 0; n < BNW / 8; ++n)
#pragma unroll
            for (int c = 0; c < 4; ++c) s[n][c] += s2[n][c];
        }
      } else {
        if (kc > 0) fetch(0, a0, b0);
#pragma unroll 1
        for (int kk = 0; kk < kc; kk += 2) {
          if (kk + 1 < kc) fetch(kk + 1, a1, b1);
          mult(a0, b0);
          if (kk + 1 < kc) {
            if (kk + 2 < kc) fetch(kk + 2, a0, b0);
            mult(a1, b1);
          }
        }
      }
      if constexpr (kSplitD) {
        // partial scores of the other shares of the features: slot [column group][row / key group], word w at
        // [w][lane]; the NC warps of a row / key group (same warp & 3) meet at their own named barrier and every one of
        // them adds the NC partials in the same order (identical scores in all of them).  The barrier that closes
        // the tile step stands between these reads and the next step's writes.
        constexpr int XW = BNW / 2;  // words per lane
        float* slot = xbuf + (warp & 3) * (XW * 32) + lane;
#pragma unroll
        for (int n = 0; n < BNW / 8; ++n)
#pragma unroll
          for (int c = 0; c < 4; ++c) slot[wc * 4 * (XW * 32) + (n * 4 + c) * 32] = s[n][c];
        asm volatile("bar.sync %0, %1;" ::"r"(1 + (warp & 3)), "n"(32 * NC) : "memory");
#pragma unroll
        for (int n = 0; n < BNW / 8; ++n)
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float a = slot[(n * 4 + c) * 32];
#pragma unroll
            for (int w = 1; w < NC; ++w) a += slot[w * 4 * (XW * 32) + (n * 4 + c) * 32];
            s[n][c] = a;
          }
      }
    }

    // ---- scale, mask, online softmax (rows g and g + 8; a row's columns live in the 4 lanes of a quad)
    // (the scores pass through the array dtype before and after the additive mask, as in the reference's op chain)
    const int j0 = tile * BN + wk * BNW;
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      float tmax = -INFINITY;
#pragma unroll
      for (int n = 0; n < BNW / 8; ++n) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int j = j0 + n * 8 + 2 * t + u;
          float x = rnd<T>(s[n][2 * e + u]);
          if (j >= rjmax[e]) {
            x = -INFINITY;
          } else if (p.mask_mode == MASK_BOOL) {
            if (rok[e] && !((const uint8_t*)p.mask)[mrow[e] + j * p.ms[3]]) x = fill;
          } else if (p.mask_mode == MASK_ADD) {
            if (rok[e]) {
              const int64_t mi = mrow[e] + j * p.ms[3];
              x = rnd<T>(x + (p.mask_is_f32 ? ((const float*)p.mask)[mi] : Num<T>::to_f(((const T*)p.mask)[mi])));
            }
          }
          s[n][2 * e + u] = x;
          tmax = fmaxf(tmax, x);
        }
      }
      tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, 1));
      tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, 2));
      const float m_new = fmaxf(m_run[e], tmax);
      const float m_safe = (m_new == -INFINITY) ? 0.f : m_new;
      const float corr = exp2f((m_run[e] - m_safe) * kLog2e);
      float psum = 0.f;
#pragma unroll
      for (int n = 0; n < BNW / 8; ++n) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const float pv = exp2f((s[n][2 * e + u] - m_safe) * kLog2e);
          s[n][2 * e + u] = pv;
          psum += pv;
        }
      }
      l_run[e] = l_run[e] * corr + psum;
      m_run[e] = m_new;
      if (__any_sync(0xffffffffu, corr != 1.0f)) {  // (rarely after the first tiles)
#pragma unroll
        for (int n = 0; n < WN / 8; ++n) {
          o[n][2 * e] *= corr;
          o[n][2 * e + 1] *= corr;
        }
      }
    }

    // ---- O += P V over this warp's WN output columns (the next V fragment is requested ahead of the MMAs)
    {
      constexpr int NKS = BNW / 16, NPV = WN / 16, NSTEP = NKS * NPV;
      uint32_t pa[NKS][4];
#pragma unroll
      for (int ks = 0; ks < NKS; ++ks) {
        pa[ks][0] = pack2<T>(s[2 * ks][0], s[2 * ks][1]);
        pa[ks][1] = pack2<T>(s[2 * ks][2], s[2 * ks][3]);
        pa[ks][2] = pack2<T>(s[2 * ks + 1][0], s[2 * ks + 1][1]);
        pa[ks][3] = pack2<T>(s[2 * ks + 1][2], s[2 * ks + 1][3]);
      }
      uint32_t vb[2][4];
      ldsm4_t(v_base, vb[0]);
#pragma unroll
      for (int i = 0; i < NSTEP; ++i) {
        const int ks = i / NPV, np = i - ks * NPV;
        if (i + 1 < NSTEP) {
          const int ks1 = (i + 1) / NPV, np1 = (i + 1) - ks1 * NPV;
          ldsm4_t(v_base + (ks1 * 16 * VP + np1 * 16) * 2, vb[(i + 1) & 1]);
        }
        mma16816<T>(o[2 * np], pa[ks], vb[i & 1][0], vb[i & 1][1]);
        mma16816<T>(o[2 * np + 1], pa[ks], vb[i & 1][2], vb[i & 1][3]);
      }
    }
    }  // warp_rows
    __syncthreads();  // the stage is free for the load issued at the top of the next step
  }
  cp_async_wait<0>();

  // ---- epilogue
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    l_run[e] += __shfl_xor_sync(0xffffffffu, l_run[e], 1);
    l_run[e] += __shfl_xor_sync(0xffffffffu, l_run[e], 2);
  }
  if constexpr (KS > 1) {
    // key groups 1 .. KS - 1 hand their states over through the (now idle) K / V stages: word w of warp slot `si` at
    // [si][w][lane], so both sides touch consecutive addresses; key group 0 folds them in
    constexpr int NW = WN / 2 + 4;  // floats per lane: O fragment + m, l of both rows
    float* xch = reinterpret_cast<float*>(Ks) + (size_t)(wc * RG + wr) * NW * 32 + lane;
    constexpr size_t kSlot = (size_t)NC * RG * NW * 32;  // floats per key group
    static_assert((KS - 1) * kSlot * 4 <= NS * (size_t)BN * (KP + VP) * 2, "hand-over area");
    __syncthreads();  // every warp is done with the stages
    if (wk > 0) {
      float* mine = xch + (size_t)(wk - 1) * kSlot;
#pragma unroll
      for (int n = 0; n < WN / 8; ++n)
#pragma unroll
        for (int c = 0; c < 4; ++c) mine[(n * 4 + c) * 32] = o[n][c];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        mine[(WN / 2 + e) * 32] = m_run[e];
        mine[(WN / 2 + 2 + e) * 32] = l_run[e];
      }
    }
    __syncthreads();
    if (wk > 0) return;
#pragma unroll 1
    for (int j = 1; j < KS; ++j) {
      const float* theirs = xch + (size_t)(j - 1) * kSlot;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float m1 = theirs[(WN / 2 + e) * 32], l1 = theirs[(WN / 2 + 2 + e) * 32];
        const float mm = fmaxf(m_run[e], m1);
        const float ms = (mm == -INFINITY) ? 0.f : mm;
        const float a0 = exp2f((m_run[e] - ms) * kLog2e), a1 = exp2f((m1 - ms) * kLog2e);
        m_run[e] = mm;
        l_run[e] = l_run[e] * a0 + l1 * a1;
#pragma unroll
        for (int n = 0; n < WN / 8; ++n) {
          o[n][2 * e] = o[n][2 * e] * a0 + theirs[(n * 4 + 2 * e) * 32] * a1;
          o[n][2 * e + 1] = o[n][2 * e + 1] * a0 + theirs[(n * 4 + 2 * e + 1) * 32] * a1;
        }
      }
    }
  }
  if (p.nsplit == 1 && !dyn) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      if (!rok[e]) continue;
      const int r = m0 + wr * 16 + g + 8 * e;
      const int hg = r - rtok[e] * p.G;
      const float inv = 1.0f / l_run[e];
      const int64_t oo = b * p.os[0] + (int64_t)(hk * p.G + hg) * p.os[1] + (int64_t)rtok[e] * p.os[2];
#pragma unroll
      for (int n = 0; n < WN / 8; ++n) {
        const int col = wc * WN + n * 8 + 2 * t;
        if (col >= p.Dv) continue;
        const float x0 = o[n][2 * e] * inv, x1 = o[n][2 * e + 1] * inv;
        if (p.out_is_f32) *reinterpret_cast<float2*>((float*)p.out + oo + col) = make_float2(x0, x1);
        else *reinterpret_cast<uint32_t*>((T*)p.out + oo + col) = pack2<T>(x0, x1);
      }
    }
  } else {
    float* part = p.part + ((((int64_t)b * p.Hkv + hk) * p.MT + mt) * p.nsplit + split) * (kBM * DVP + 2 * kBM);
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      if (!rok[e]) continue;
      const int row = wr * 16 + g + 8 * e;
#pragma unroll
      for (int n = 0; n < WN / 8; ++n) {
        const int col = wc * WN + n * 8 + 2 * t;
        if (col >= p.Dv) continue;
        *reinterpret_cast<float2*>(part + row * DVP + col) = make_float2(o[n][2 * e], o[n][2 * e + 1]);
      }
      if (wc == 0 && t == 0) {
        part[kBM * DVP + row] = m_run[e];
        part[kBM * DVP + kBM + row] = l_run[e];
      }
    }
  }
}

// Merge of the split partials: one CTA of 512 threads per packed query row.  The weights of the splits are settled
// once in shared memory; the threads then cover (float4 column) x (split group) so that the loads of the fold are
// independent and few per thread (a row's partials are nsplit x Dv floats spread over L2).
constexpr int kCT = 512;
template <typename T>
__global__ void __launch_bounds__(kCT) sdpa_mma_combine_kernel(const __grid_constant__ MmaParams p, int DVP) {
  __shared__ float ws[160];
  __shared__ float red[2 * kCT / 32];
  __shared__ float4 acc4[kCT];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int r = blockIdx.x, hk = blockIdx.y, b = blockIdx.z;
  const int mt = r / kBM, row = r - mt * kBM;
  const int64_t pstride = (int64_t)kBM * DVP + 2 * kBM;
  const float* part = p.part + (((int64_t)b * p.Hkv + hk) * p.MT + mt) * p.nsplit * pstride;
  // nsplit <= 148 (launch_cfg): one split per thread
  const float m0 = tid < p.nsplit ? part[tid * pstride + kBM * DVP + row] : -INFINITY;
  const float l0 = tid < p.nsplit ? part[tid * pstride + kBM * DVP + kBM + row] : 0.f;
  float M = warp_max(m0);
  if (lane == 0) red[warp] = M;
  __syncthreads();
  M = red[0];
#pragma unroll
  for (int i = 1; i < 5; ++i) M = fmaxf(M, red[i]);  // warps 0 .. 4 hold the splits
  if (M == -INFINITY) return;  // no key anywhere (paged mode: a released slot): the output row is left alone
  const float m_safe = M;
  const float w0 = exp2f((m0 - m_safe) * kLog2e);
  if (tid < p.nsplit) ws[tid] = w0;
  const float lsum = warp_sum(w0 * l0);
  if (lane == 0) red[kCT / 32 + warp] = lsum;
  __syncthreads();
  float L = 0.f;
#pragma unroll
  for (int i = 0; i < 5; ++i) L += red[kCT / 32 + i];
  const float inv = 1.0f / L;

  const int ncol4 = p.Dv >> 2;   // <= 128 (values up to 512 features, multiples of 8)
  const int nsg = kCT / ncol4;   // split groups
  const int sg = tid / ncol4, c4 = tid - sg * ncol4;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  if (sg < nsg) {
    const float* src = part + row * DVP + 4 * c4;
#pragma unroll 8
    for (int s = sg; s < p.nsplit; s += nsg) {
      const float w = ws[s];
      if (w == 0.f) continue;  // (no visible key in that split, or -- graph mode -- no such split: nothing stored)
      const float4 v = *reinterpret_cast<const float4*>(src + s * pstride);
      a.x = fmaf(w, v.x, a.x);
      a.y = fmaf(w, v.y, a.y);
      a.z = fmaf(w, v.z, a.z);
      a.w = fmaf(w, v.w, a.w);
    }
  }
  acc4[tid] = a;
  __syncthreads();
  if (sg != 0) return;
  for (int g2 = 1; g2 < nsg; ++g2) {
    const float4 v = acc4[g2 * ncol4 + c4];
    a.x += v.x;
    a.y += v.y;
    a.z += v.z;
    a.w += v.w;
  }
  const int tok = r / p.G, hg = r - tok * p.G;
  const int64_t oo = b * p.os[0] + (int64_t)(hk * p.G + hg) * p.os[1] + (int64_t)tok * p.os[2] + 4 * c4;
  if (p.out_is_f32) {
    *reinterpret_cast<float2*>((float*)p.out + oo) = make_float2(a.x * inv, a.y * inv);
    *reinterpret_cast<float2*>((float*)p.out + oo + 2) = make_float2(a.z * inv, a.w * inv);
  } else {
    *reinterpret_cast<uint32_t*>((T*)p.out + oo) = pack2<T>(a.x * inv, a.y * inv);
    *reinterpret_cast<uint32_t*>((T*)p.out + oo + 2) = pack2<T>(a.z * inv, a.w * inv);
  }
}

bool rows16h(const omx_array* a) {  // feature axis contiguous, rows 16-byte aligned (8 two-byte elements)
  if (a->strides[3] != 1 && a->shape[3] != 1) return false;
  if (reinterpret_cast<uintptr_t>(a->data) & 15) return false;
  for (int i = 0; i < 3; ++i)
    if (a->shape[i] > 1 && a->strides[i] % 8) return false;
  return true;
}

template <typename T, int DKP, int DVP, int KS = 1>
void launch_cfg(MmaParams& p, cudaStream_t stream) {
  using C = MmaCfg<DKP, DVP, KS>;
  auto kern = sdpa_mma_kernel<T, DKP, DVP, KS>;
  constexpr size_t smem = C::smem;
  OMX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // split the key range when the row tiles alone leave SMs idle: at least 2 key tiles (both stages) per split
  const int64_t base = (int64_t)p.MT * p.Hkv * p.B;
  static const int min_tiles = [] {
    const char* e = getenv("OMX_MMA_MIN_TILES");
    return e ? std::max(1, atoi(e)) : 2;
  }();
  // as many CTAs as stay resident (the combine kernel handles up to 148 splits)
  const int64_t slots = (int64_t)sm_count() * C::MINB;
  p.split_cap = base < slots ? (int)std::min<int64_t>(slots / base, 148) : 1;
  p.min_tiles = min_tiles;
  const int nsplit = mma_plan_splits(p.Lk, C::BN, p.split_cap, p.min_tiles);
  p.nsplit = nsplit;
  const size_t part_bytes = sizeof(float) * (size_t)base * nsplit * (kBM * DVP + 2 * kBM);
  const bool dyn = p.pos_dev != nullptr || p.lens_in != nullptr;
  const bool two_launches = nsplit > 1 || dyn;
  if (dyn) {  // graph / paged mode: partials live in the caller's (cache-owned, fixed-address) scratch
    OMX_CHECK(p.part && part_bytes <= p.part_bytes, "[sdpa_mma] graph-mode scratch too small (%zu > %zu bytes)",
              part_bytes, p.part_bytes);
  } else if (nsplit > 1) {
    p.part = (float*)get_workspace(part_bytes, stream);
  }
  dim3 grid(p.MT * nsplit, p.Hkv, p.B);
  kern<<<grid, C::NT, smem, stream>>>(p);
  count_launch();
  OMX_CUDA(cudaGetLastError());
  if (two_launches) {
    dim3 cgrid(p.R, p.Hkv, p.B);
    sdpa_mma_combine_kernel<T><<<cgrid, kCT, 0, stream>>>(p, DVP);
    count_launch();
    OMX_CUDA(cudaGetLastError());
  }
}

template <typename T, int DKP, int DVP>
void launch_ks(MmaParams& p, cudaStream_t stream) {
  // decode-sized row counts: key groups instead of idle row groups (OMX_MMA_NO_KS=1: A/B runs)
  const bool ks_ok = !getenv("OMX_MMA_NO_KS");
  if constexpr (DKP <= 256) {
    if (ks_ok && p.R <= 16) return launch_cfg<T, DKP, DVP, 4>(p, stream);
  }
  if (ks_ok && p.R <= 32) return launch_cfg<T, DKP, DVP, 2>(p, stream);
  launch_cfg<T, DKP, DVP, 1>(p, stream);
}

template <typename T>
void launch_t(MmaParams& p, cudaStream_t stream) {
  const int d = p.D, dv = p.Dv;
  if (d <= 32 && dv <= 32) launch_ks<T, 32, 32>(p, stream);
  else if (d <= 64 && dv <= 64) launch_ks<T, 64, 64>(p, stream);
  else if (d <= 80 && dv <= 80) launch_ks<T, 80, 80>(p, stream);  // vision towers: head dim 80 (and 72, zero-padded)
  else if (d <= 96 && dv <= 96) launch_ks<T, 96, 96>(p, stream);
  else if (d <= 128 && dv <= 128) launch_ks<T, 128, 128>(p, stream);
  else if (d <= 256 && dv <= 256) launch_ks<T, 256, 256>(p, stream);
  else launch_ks<T, 576, 512>(p, stream);
}

}  // namespace

bool sdpa_mma_supported(const SdpaArgs& a, const char** why) {
  auto no = [&](const char* w) {
    if (why) *why = w;
    return false;
  };
  const int dt = a.q->dtype;
  if (!(dt == OMX_BFLOAT16 || dt == OMX_FLOAT16) || a.k->dtype != dt || a.v->dtype != dt) return no("not bf16 / f16");
  if (a.out->dtype != dt && a.out->dtype != OMX_FLOAT32) return no("output dtype");
  if (a.D % 8 || a.Dv % 8 || a.D > 576 || a.Dv > 512 || a.D < 8 || a.Dv < 8)
    return no("head dims must be multiples of 8, keys <= 576, values <= 512");
  if (a.Lk < 1) return no("no keys");
  if (a.Hkv < 1 || a.Hq % a.Hkv) return no("query heads not a multiple of kv heads");
  if (a.Hkv > 65535 || a.B > 65535) return no("grid too large");
  if (a.mask_mode == MASK_ADD && !(a.mask->dtype == dt || a.mask->dtype == OMX_FLOAT32)) return no("additive mask dtype");
  if (!rows16h(a.q) || !rows16h(a.k) || !rows16h(a.v)) return no("rows not contiguous / 16-byte aligned");
  // the output is stored as column pairs
  if ((a.out->strides[3] != 1 && a.out->shape[3] != 1) ||
      (reinterpret_cast<uintptr_t>(a.out->data) & 7))
    return no("output rows not contiguous / aligned");
  for (int i = 0; i < 3; ++i)
    if (a.out->shape[i] > 1 && a.out->strides[i] % 2) return no("output rows not aligned");
  return true;
}

// Single-token calls the TMA decode kernel does not take (16-bit, head dim != 128) used to run on the CUDA-core
// split-K kernel.  With two or more query heads per kv head the key-group variants here are 2 - 7x faster (B32,
// ctx 4096, bf16: head dim 256, 16 / 2 heads 260 -> 72 us; 64, 16 / 4 heads 210 -> 62 us; 32: 202 -> 29 us --
// scripts/gpu_r02_simt_vs_mma.py); one query head per kv head stays there (182 vs 258 us at head dim 256).
bool sdpa_mma_preferred_for_decode(const SdpaArgs& a) {
  const int dt = a.q->dtype;
  if (!(dt == OMX_BFLOAT16 || dt == OMX_FLOAT16) || a.Lq != 1) return false;
  if (a.D == 128 && a.Dv == 128) return false;  // decode_hmma_tma's shape
  // one query head per kv head: narrow heads still win here (B32, 16 heads, d64, ctx 4096: 245 -> 146 us), 256-wide
  // ones do not (182 vs 258 us)
  if (a.Hkv < 1 || (a.Hq / a.Hkv < 2 && (a.D > 64 || a.Dv > 64))) return false;
  return sdpa_mma_supported(a, nullptr);
}

size_t sdpa_mma_graph_scratch_bytes(int B, int Hkv, int Hq, int Dv) {
  // worst case over the width configurations: every CTA slot a split, value rows padded to the configuration's width
  const int dvp = Dv <= 32 ? 32 : Dv <= 64 ? 64 : Dv <= 80 ? 80 : Dv <= 96 ? 96 : Dv <= 128 ? 128 : Dv <= 256 ? 256 : 512;
  const int64_t base = (int64_t)B * Hkv * ((Hq / std::max(Hkv, 1) + kBM - 1) / kBM);
  const int64_t slots = (int64_t)sm_count() * 3;
  const int64_t nsplit = base < slots ? std::min<int64_t>(slots / base, 148) : 1;
  return sizeof(float) * (size_t)base * (size_t)nsplit * (kBM * dvp + 2 * kBM);
}

void sdpa_mma(const SdpaArgs& a, cudaStream_t stream) { sdpa_mma_dynamic(a, nullptr, nullptr, 0, stream); }

void sdpa_mma_dynamic(const SdpaArgs& a, const int* pos_dev, void* scratch, size_t scratch_bytes, cudaStream_t stream,
                      const PagedRef* paged) {
  MmaParams p{};
  p.pos_dev = pos_dev;
  p.part_bytes = scratch_bytes;
  if (paged) {
    p.lens_in = paged->lens_in;
    p.block_table = paged->block_table;
    p.bt_stride = paged->bt_stride;
  }
  p.q = a.q->data;
  p.k = a.k->data;
  p.v = a.v->data;
  p.out = a.out->data;
  p.mask = a.mask ? a.mask->data : nullptr;
  p.part = (float*)scratch;
  for (int i = 0; i < 3; ++i) {
    p.qs[i] = a.q->strides[i];
    p.ks[i] = a.k->strides[i];
    p.vs[i] = a.v->strides[i];
    p.os[i] = a.out->strides[i];
  }
  for (int i = 0; i < 4; ++i) p.ms[i] = a.mask_strides[i];
  p.B = a.B; p.Hq = a.Hq; p.Hkv = a.Hkv; p.G = a.Hq / a.Hkv; p.Lq = a.Lq; p.Lk = a.Lk; p.D = a.D; p.Dv = a.Dv;
  p.R = p.G * p.Lq;
  p.MT = (p.R + kBM - 1) / kBM;
  p.scale = a.scale;
  p.mask_mode = a.mask_mode;
  p.mask_is_f32 = (a.mask && a.mask->dtype == OMX_FLOAT32) ? 1 : 0;
  p.out_is_f32 = a.out->dtype == OMX_FLOAT32 ? 1 : 0;
  note_launch("sdpa_mma");
  if (a.q->dtype == OMX_BFLOAT16) launch_t<__nv_bfloat16>(p, stream);
  else launch_t<__half>(p, stream);
}

}  // namespace omx
