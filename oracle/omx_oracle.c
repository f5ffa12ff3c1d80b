/*
 * omx_oracle.c -- CPU restatement of the reference's attention hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ may be imported, linked or
 * executed by the product path (ominix-mlx_b200/, include/).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * use it, and only as the checker or as the timed CPU baseline.
 *
 * What it restates.  The arithmetic of this path is NOT inside /root/reference:
 * mlx-rs/src/fast.rs:15-46 (rope) and :121-151 (scaled_dot_product_attention)
 * are one-call FFI shims into ml-explore/mlx v0.30.1 (pinned by
 * mlx-rs/mlx-sys/src/mlx-c/CMakeLists.txt:35-38, reached through
 * mlx-c/mlx/c/fast.cpp:557 and :617).  On the MLX *CPU* backend both ops run
 * as their published "fallback" primitive graphs; this file restates those
 * graphs op by op, with a rounding to the array dtype after every primitive,
 * exactly as a chain of separate MLX CPU kernels produces.
 *
 * Pinning.  rope is pinned against the reference's only golden vector for this
 * path (mlx-rs/src/fast.rs:231-251 == mlx-rs/src/nn/positional_encoding.rs:
 * 432-463: seed 71, uniform [2,8,16], mean 0.45625377 / sum 116.80096) through
 * the threefry restatement in oracle/mlx_random.py -- see
 * tests/test_oracle_golden.py.  sdpa and the KV cache have NO value-level test
 * in the reference (fast.rs:301-331 checks shape/dtype only; cache.rs has no
 * tests): they are cross-checked against torch CPU fp64 attention and against
 * the hand-derived cases of SURVEY.md Appendix A instead.
 *
 * Build: see oracle/Makefile.  -ffp-contract=off is REQUIRED: MLX evaluates
 * multiply and subtract as separate kernels, so no FMA contraction may happen.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* dtype codes = mlx_dtype (mlx-c/mlx/c/array.h:37-52) */
enum { OMX_BOOL = 0, OMX_F16 = 9, OMX_F32 = 10, OMX_BF16 = 12 };

/* ---- scalar dtype helpers ------------------------------------------------ */

static inline float bf16_to_f32(uint16_t h) {
  uint32_t u = ((uint32_t)h) << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

/* round-to-nearest-even, NaN preserved (same as MLX's bfloat16_t conversion) */
static inline uint16_t f32_to_bf16(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x0040u);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

static inline float f16_to_f32(uint16_t h) {
  _Float16 x;
  memcpy(&x, &h, 2);
  return (float)x;
}

static inline uint16_t f32_to_f16(float f) {
  _Float16 x = (_Float16)f;
  uint16_t h;
  memcpy(&h, &x, 2);
  return h;
}

static inline size_t dt_size(int dt) { return dt == OMX_F32 ? 4 : 2; }

static inline float ld(const void* p, int dt, size_t i) {
  if (dt == OMX_F32) return ((const float*)p)[i];
  if (dt == OMX_BF16) return bf16_to_f32(((const uint16_t*)p)[i]);
  return f16_to_f32(((const uint16_t*)p)[i]);
}

static inline void st(void* p, int dt, size_t i, float v) {
  if (dt == OMX_F32) ((float*)p)[i] = v;
  else if (dt == OMX_BF16) ((uint16_t*)p)[i] = f32_to_bf16(v);
  else ((uint16_t*)p)[i] = f32_to_f16(v);
}

/* value after being stored in an array of dtype dt */
static inline float rnd(float v, int dt) {
  if (dt == OMX_F32) return v;
  if (dt == OMX_BF16) return bf16_to_f32(f32_to_bf16(v));
  return f16_to_f32(f32_to_f16(v));
}

static inline float dt_lowest(int dt) { /* finfo(dtype).min */
  if (dt == OMX_F16) return -65504.0f;
  if (dt == OMX_BF16) return -3.3895313892515355e38f;
  return -3.4028234663852886e38f;
}

int omx_oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void omx_oracle_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ---- rope ---------------------------------------------------------------- */

/*
 * cos/sin table of the rope fallback graph (MLX v0.30.1 mlx/fast.cpp, rope()
 * fallback lambda; reached from mlx-c fast.cpp:557):
 *   positions = (arange(T, f32) + offset) * scale          (f32 ops)
 *   inv_freqs = exp(arange(0, -half, -1, f32) * (log(base) / half))
 *             | reciprocal(freqs)
 *   theta     = positions[:, None] * inv_freqs             (ONE f32 multiply)
 *   cos/sin(theta) evaluated in f32 (std::cos/std::sin on x86: Simd<T,1>)
 * cos_out/sin_out: [T, half] f32 (before the cast to x.dtype).
 */
void omx_oracle_rope_table(float* cos_out, float* sin_out, int T, int dims,
                           int has_base, float base, float scale, int offset,
                           const float* freqs) {
  int half = dims / 2;
  float* inv = (float*)malloc(sizeof(float) * (size_t)(half > 0 ? half : 1));
  if (freqs) {
    for (int i = 0; i < half; ++i) inv[i] = 1.0f / freqs[i];
  } else {
    (void)has_base;
    float step = logf(base) / (float)half;
    for (int i = 0; i < half; ++i) inv[i] = expf((float)(-i) * step);
  }
  for (int t = 0; t < T; ++t) {
    float pos = ((float)t + (float)offset) * scale;
    for (int i = 0; i < half; ++i) {
      float th = pos * inv[i];
      cos_out[(size_t)t * half + i] = cosf(th);
      sin_out[(size_t)t * half + i] = sinf(th);
    }
  }
  free(inv);
}

/*
 * fast::rope on a contiguous [B, N, T, D] array (mlx-rs/src/fast.rs:15-46).
 * All arithmetic after the table is done in x's dtype, rounding after every
 * primitive:  o1 = x1*c - x2*s ; o2 = x1*s + x2*c  (forward rotation).
 *   traditional=0: x1 = x[..., :half], x2 = x[..., half:dims]   (halves)
 *   traditional=1: x1 = x[..., 0:dims:2], x2 = x[..., 1:dims:2] (interleaved)
 * Features >= dims are copied.  Same position for every batch row.
 */
void omx_oracle_rope(const void* x, void* out, int dt, int B, int N, int T,
                     int D, int dims, int traditional, int has_base,
                     float base, float scale, int offset, const float* freqs) {
  int half = dims / 2;
  size_t tbl = (size_t)T * (size_t)(half > 0 ? half : 1);
  float* c = (float*)malloc(sizeof(float) * tbl);
  float* s = (float*)malloc(sizeof(float) * tbl);
  omx_oracle_rope_table(c, s, T, dims, has_base, base, scale, offset, freqs);
  for (size_t i = 0; i < (size_t)T * half; ++i) { /* astype(cos(theta), t) */
    c[i] = rnd(c[i], dt);
    s[i] = rnd(s[i], dt);
  }
  long rows = (long)B * N * T;
#pragma omp parallel for schedule(static)
  for (long r = 0; r < rows; ++r) {
    int t = (int)(r % T);
    size_t o = (size_t)r * D;
    const float* ct = c + (size_t)t * half;
    const float* stt = s + (size_t)t * half;
    for (int i = 0; i < half; ++i) {
      size_t i1 = traditional ? (size_t)(2 * i) : (size_t)i;
      size_t i2 = traditional ? (size_t)(2 * i + 1) : (size_t)(i + half);
      float x1 = ld(x, dt, o + i1), x2 = ld(x, dt, o + i2);
      float a = rnd(x1 * ct[i], dt), b = rnd(x2 * stt[i], dt);
      float e = rnd(x1 * stt[i], dt), f = rnd(x2 * ct[i], dt);
      st(out, dt, o + i1, a - b);
      st(out, dt, o + i2, e + f);
    }
    for (int d = dims; d < D; ++d) st(out, dt, o + d, ld(x, dt, o + d));
  }
  free(c);
  free(s);
}

/* ---- rms_norm -------------------------------------------------------------- */

/*
 * fast::rms_norm CPU fallback graph (MLX v0.30.1 mlx/fast.cpp, reached from
 * mlx-c fast.cpp via mlx_fast_rms_norm, fast.h:163-168; Rust entry
 * mlx-rs/src/fast.rs:163-180; callers: q_norm / k_norm per head,
 * qwen3-mlx/src/model.rs:172-181):
 *   xf  = astype(x, float32)
 *   m   = mean(square(xf), -1, keepdims) = sum(xf*xf) * float32(1/D)
 *         (ops.cpp mean: multiply(sum, 1/n); the contiguous CPU reduce walks
 *          the row left to right -- on x86 simd::max_size is 1, base_simd.h)
 *   n   = xf * rsqrt(m + eps)          rsqrt = 1.0f / sqrtf(.) on the CPU backend
 *   y   = astype(n, dtype)
 *   out = weight * y                   (in dtype; skipped without a weight)
 * x [rows, D] contiguous, weight [D] in the same dtype or NULL.
 */
void omx_oracle_rms_norm(const void* x, const void* w, void* out, int dt,
                         long rows, int D, float eps) {
  const float inv_n = 1.0f / (float)D;
#pragma omp parallel for schedule(static)
  for (long r = 0; r < rows; ++r) {
    size_t o = (size_t)r * D;
    float acc = 0.0f;
    for (int d = 0; d < D; ++d) {
      float v = ld(x, dt, o + d);
      float sq = v * v;
      acc = acc + sq;
    }
    float m = acc * inv_n;
    float rs = 1.0f / sqrtf(m + eps);
    for (int d = 0; d < D; ++d) {
      float y = rnd(ld(x, dt, o + d) * rs, dt);
      if (w) y = ld(w, dt, (size_t)d) * y;
      st(out, dt, o + d, y);
    }
  }
}

/* ---- scaled_dot_product_attention ---------------------------------------- */

enum { MASK_NONE = 0, MASK_CAUSAL = 1, MASK_BOOL = 2, MASK_ADD = 3 };

/*
 * fast::scaled_dot_product_attention CPU fallback graph (MLX v0.30.1
 * mlx/fast.cpp, reached from mlx-c fast.cpp:617; Rust entry
 * mlx-rs/src/fast.rs:121-151), contiguous inputs:
 *   q [B,Hq,Lq,D]  k [B,Hkv,Lk,D]  v [B,Hkv,Lk,Dv]  out [B,Hq,Lq,Dv]
 *   q'     = array(scale, dtype) * q                       (rounded to dtype)
 *   scores = matmul(q', k^T)         f32 accumulate, rounded to dtype
 *   causal : keep iff (max(Lk-Lq,0) + i) >= j
 *   bool   : where(mask, scores, finfo(dtype).min)   [neg_inf!=0: -inf]
 *   float  : scores + mask                                 (rounded to dtype)
 *   p      = softmax(scores, precise=true): f32 max / exp / sum,
 *            exp * (1/sum) rounded to dtype
 *   out    = matmul(p, v)            f32 accumulate, rounded to dtype
 * GQA: q head h reads kv head h / (Hq/Hkv)  (fast.rs:118: K/V not pre-tiled).
 * mask: element strides ms[4] of the mask broadcast to [B,Hq,Lq,Lk] (0 on
 * broadcast axes); mask_dt = OMX_BOOL (uint8) or a float dtype.
 * softmax_in_dtype != 0 restates softmax(precise=false): arithmetic in dtype.
 */
void omx_oracle_sdpa(const void* q, const void* k, const void* v, void* out,
                     int dt, int B, int Hq, int Hkv, int Lq, int Lk, int D,
                     int Dv, float scale, int mask_mode, const void* mask,
                     int mask_dt, const int64_t* ms, int bool_fill_neg_inf) {
  int G = Hq / Hkv;
  float scale_t = rnd(scale, dt);
  int q_off = (Lk - Lq) < 0 ? 0 : (Lk - Lq);
  float fill = bool_fill_neg_inf ? -INFINITY : dt_lowest(dt);
  long rows = (long)B * Hq * Lq;
#pragma omp parallel
  {
    float* qs = (float*)malloc(sizeof(float) * (size_t)D);
    float* sc = (float*)malloc(sizeof(float) * (size_t)Lk);
    float* acc = (float*)malloc(sizeof(float) * (size_t)Dv);
#pragma omp for schedule(dynamic, 1)
    for (long r = 0; r < rows; ++r) {
      int i = (int)(r % Lq);
      int h = (int)((r / Lq) % Hq);
      int b = (int)(r / ((long)Lq * Hq));
      int hk = h / G;
      size_t qo = (size_t)r * D;
      size_t ko = ((size_t)b * Hkv + hk) * (size_t)Lk * D;
      size_t vo = ((size_t)b * Hkv + hk) * (size_t)Lk * Dv;
      for (int d = 0; d < D; ++d) qs[d] = rnd(scale_t * ld(q, dt, qo + d), dt);
      for (int j = 0; j < Lk; ++j) {
        float a = 0.f;
        size_t kr = ko + (size_t)j * D;
        if (dt == OMX_F32) {
          const float* kp = (const float*)k + kr;
          for (int d = 0; d < D; ++d) a += qs[d] * kp[d];
        } else if (dt == OMX_BF16) {
          const uint16_t* kp = (const uint16_t*)k + kr;
          for (int d = 0; d < D; ++d) a += qs[d] * bf16_to_f32(kp[d]);
        } else {
          for (int d = 0; d < D; ++d) a += qs[d] * ld(k, dt, kr + d);
        }
        a = rnd(a, dt);
        if (mask_mode == MASK_CAUSAL) {
          if (!(q_off + i >= j)) a = fill;
        } else if (mask_mode == MASK_BOOL) {
          size_t mi = (size_t)(b * ms[0] + h * ms[1] + i * ms[2] + j * ms[3]);
          if (!((const uint8_t*)mask)[mi]) a = fill;
        } else if (mask_mode == MASK_ADD) {
          size_t mi = (size_t)(b * ms[0] + h * ms[1] + i * ms[2] + j * ms[3]);
          a = rnd(a + ld(mask, mask_dt, mi), dt);
        }
        sc[j] = a;
      }
      float mx = -INFINITY;
      for (int j = 0; j < Lk; ++j) mx = sc[j] > mx ? sc[j] : mx;
      float sum = 0.f;
      for (int j = 0; j < Lk; ++j) {
        sc[j] = expf(sc[j] - mx);
        sum += sc[j];
      }
      float inv = 1.0f / sum;
      for (int d = 0; d < Dv; ++d) acc[d] = 0.f;
      for (int j = 0; j < Lk; ++j) {
        float p = rnd(sc[j] * inv, dt);
        size_t vr = vo + (size_t)j * Dv;
        if (dt == OMX_F32) {
          const float* vp = (const float*)v + vr;
          for (int d = 0; d < Dv; ++d) acc[d] += p * vp[d];
        } else if (dt == OMX_BF16) {
          const uint16_t* vp = (const uint16_t*)v + vr;
          for (int d = 0; d < Dv; ++d) acc[d] += p * bf16_to_f32(vp[d]);
        } else {
          for (int d = 0; d < Dv; ++d) acc[d] += p * ld(v, dt, vr + d);
        }
      }
      size_t oo = (size_t)r * Dv;
      for (int d = 0; d < Dv; ++d) st(out, dt, oo + d, acc[d]);
    }
    free(qs);
    free(sc);
    free(acc);
  }
}

/*
 * DiT joint attention, the MANUAL op chain of the image crates (NOT fast::sdpa):
 *   FLUX.2-klein flux-klein-mlx/src/klein_model.rs:474-483 and :651-659:
 *     attn = matmul(q, k^T)            (dtype t, f32 accumulate)
 *     attn = attn / array!(sqrt(D))    (f32 scalar ARRAY => result promotes to f32)
 *     attn = softmax_axis(attn, -1, precise=None)   (f32 by then)
 *     out  = matmul(attn, v)           (f32 x t => f32)
 *   Z-Image zimage-mlx/src/zimage_model.rs:368-384: same with `* array!(scale)`
 *     and an optional additive mask (use_mul != 0; add_mask f32 [Lq,Lk] or NULL).
 * q [B,H,Lq,D], k/v [B,H,Lk,D] contiguous (GQA already repeated by the caller,
 * zimage_model.rs:360-367).  out is ALWAYS f32 [B,H,Lq,D] for 16-bit inputs too.
 */
void omx_oracle_dit_attention(const void* q, const void* k, const void* v,
                              float* out, int dt, int B, int H, int Lq, int Lk,
                              int D, float scale_or_div, int use_mul,
                              const float* add_mask) {
  long rows = (long)B * H * Lq;
#pragma omp parallel
  {
    float* sc = (float*)malloc(sizeof(float) * (size_t)Lk);
    float* acc = (float*)malloc(sizeof(float) * (size_t)D);
    float* qf = (float*)malloc(sizeof(float) * (size_t)D);
#pragma omp for schedule(dynamic, 1)
    for (long r = 0; r < rows; ++r) {
      int i = (int)(r % Lq);
      size_t bh = (size_t)(r / Lq);
      size_t qo = (size_t)r * D, ko = bh * (size_t)Lk * D;
      for (int d = 0; d < D; ++d) qf[d] = ld(q, dt, qo + d);
      for (int j = 0; j < Lk; ++j) {
        float a = 0.f;
        size_t kr = ko + (size_t)j * D;
        if (dt == OMX_BF16) {
          const uint16_t* kp = (const uint16_t*)k + kr;
          for (int d = 0; d < D; ++d) a += qf[d] * bf16_to_f32(kp[d]);
        } else {
          for (int d = 0; d < D; ++d) a += qf[d] * ld(k, dt, kr + d);
        }
        a = rnd(a, dt); /* matmul output in t */
        a = use_mul ? a * scale_or_div : a / scale_or_div; /* f32 from here */
        if (add_mask) a += add_mask[(size_t)i * Lk + j];
        sc[j] = a;
      }
      float mx = -INFINITY;
      for (int j = 0; j < Lk; ++j) mx = sc[j] > mx ? sc[j] : mx;
      float sum = 0.f;
      for (int j = 0; j < Lk; ++j) {
        sc[j] = expf(sc[j] - mx);
        sum += sc[j];
      }
      float inv = 1.0f / sum;
      for (int d = 0; d < D; ++d) acc[d] = 0.f;
      for (int j = 0; j < Lk; ++j) {
        float p = sc[j] * inv;
        size_t vr = ko + (size_t)j * D;
        if (dt == OMX_BF16) {
          const uint16_t* vp = (const uint16_t*)v + vr;
          for (int d = 0; d < D; ++d) acc[d] += p * bf16_to_f32(vp[d]);
        } else {
          for (int d = 0; d < D; ++d) acc[d] += p * ld(v, dt, vr + d);
        }
      }
      for (int d = 0; d < D; ++d) out[(size_t)r * D + d] = acc[d];
    }
    free(sc);
    free(acc);
    free(qf);
  }
}

/*
 * DiT table-driven interleaved rope (flux-klein-mlx/src/klein_model.rs:124-162,
 * zimage-mlx/src/zimage_model.rs:208-235): x [B,S,H,D], cos/sin [B,S,D/2] in
 * x's dtype (klein keeps a duplicated [B,S,D] table and reads element 0 of each
 * pair -- same values).  out0 = x0*c - x1*s ; out1 = x1*c + x0*s (klein order);
 * zimage writes out1 = x0*s + x1*c -- the same sum, addition commutes.
 */
void omx_oracle_dit_rope(const void* x, const void* cs, const void* sn,
                         void* out, int dt, int B, int S, int H, int D) {
  int half = D / 2;
  long rows = (long)B * S * H;
#pragma omp parallel for schedule(static)
  for (long r = 0; r < rows; ++r) {
    size_t bs = (size_t)(r / H);
    size_t o = (size_t)r * D, to = bs * half;
    for (int i = 0; i < half; ++i) {
      float x0 = ld(x, dt, o + 2 * i), x1 = ld(x, dt, o + 2 * i + 1);
      float c = ld(cs, dt, to + i), s = ld(sn, dt, to + i);
      float a = rnd(x0 * c, dt), b = rnd(x1 * s, dt);
      float e = rnd(x1 * c, dt), f = rnd(x0 * s, dt);
      st(out, dt, o + 2 * i, a - b);
      st(out, dt, o + 2 * i + 1, e + f);
    }
  }
}
