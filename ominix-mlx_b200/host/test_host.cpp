// test_host.cpp -- self-checking tests of the C++ host layer (omx_attn.hpp) on a B200.
// Mirrors what the reference's own tests exercise for this path (mlx-rs/src/fast.rs:231-331,
// mlx-rs-core/src/cache.rs worked examples of SURVEY Appendix A) plus parity against the CPU oracle
// (oracle/libomx_oracle.so -- the checker, linked by this TEST only, never by the library).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>

#include "omx_attn.hpp"

extern "C" {
void omx_oracle_rope(const void* x, void* out, int dt, int B, int N, int T, int D, int dims, int traditional,
                     int has_base, float base, float scale, int offset, const float* freqs);
void omx_oracle_rms_norm(const void* x, const void* w, void* out, int dt, long rows, int D, float eps);
void omx_oracle_sdpa(const void* q, const void* k, const void* v, void* out, int dt, int B, int Hq, int Hkv, int Lq,
                     int Lk, int D, int Dv, float scale, int mask_mode, const void* mask, int mask_dt,
                     const int64_t* ms, int bool_fill_neg_inf);
}

using omx::Array;
using omx::Dtype;

static int g_failed = 0;
#define EXPECT(cond, ...)                                   \
  do {                                                      \
    if (!(cond)) {                                          \
      ++g_failed;                                           \
      printf("  FAILED %s:%d: ", __FILE__, __LINE__);       \
      printf(__VA_ARGS__);                                  \
      printf("\n");                                         \
    }                                                       \
  } while (0)

static uint16_t f2bf(float f) {  // round to nearest even
  uint32_t u;
  memcpy(&u, &f, 4);
  u += 0x7FFF + ((u >> 16) & 1);
  return (uint16_t)(u >> 16);
}
static float bf2f(uint16_t b) {
  uint32_t u = (uint32_t)b << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}
static std::vector<uint16_t> randn_bf16(size_t n, unsigned seed) {
  std::mt19937 g(seed);
  std::normal_distribution<float> d(0.f, 1.f);
  std::vector<uint16_t> v(n);
  for (auto& x : v) x = f2bf(d(g));
  return v;
}
static std::vector<uint16_t> download(const Array& a) {  // any strides -> contiguous host copy
  std::vector<uint16_t> h((size_t)a.size());
  if (a.is_contiguous()) {
    a.to_host(h.data());
    return h;
  }
  // gather through a contiguous device copy made by the library's own strided copy: rope with dims... no --
  // use cudaMemcpy2D over the innermost-contiguous rows
  const auto sh = a.shape(), st = a.strides();
  const int64_t D = sh.back();
  std::vector<int64_t> idx(sh.size() - 1, 0);
  size_t row = 0;
  const int64_t rows = a.size() / D;
  for (int64_t r = 0; r < rows; ++r, ++row) {
    int64_t off = 0;
    for (size_t i = 0; i + 1 < sh.size(); ++i) off += idx[i] * st[i];
    cudaMemcpy(h.data() + row * D, (const uint16_t*)a.data() + off, (size_t)D * 2, cudaMemcpyDeviceToHost);
    for (int i = (int)sh.size() - 2; i >= 0; --i) {
      if (++idx[i] < sh[i]) break;
      idx[i] = 0;
    }
  }
  return h;
}

static void test_rope_bit_exact() {
  printf("test_rope_bit_exact\n");
  const int B = 2, H = 4, T = 5, D = 128, offset = 1234;
  auto x = randn_bf16((size_t)B * H * T * D, 1);
  Array xa = Array::from_host(x.data(), {B, H, T, D}, Dtype::Bfloat16);
  Array out = omx::fast::rope(xa, D, false, 1e6f, 1.0f, offset);
  EXPECT(out.shape() == xa.shape() && out.dtype() == Dtype::Bfloat16, "shape/dtype");
  std::vector<uint16_t> want(x.size());
  omx_oracle_rope(x.data(), want.data(), OMX_BFLOAT16, B, H, T, D, D, 0, 1, 1e6f, 1.0f, offset, nullptr);
  EXPECT(download(out) == want, "rope output differs from the oracle bit pattern");
  // nn::Rope through the caller's transposed layout: [B, T, H, D] storage viewed [B, H, T, D]
  Array xt = Array::from_host(x.data(), {B, T, H, D}, Dtype::Bfloat16).transpose_axes({0, 2, 1, 3});
  omx::nn::Rope rope{D, false, 1e6f, 1.0f};
  Array o2 = rope.forward(xt, offset);
  std::vector<uint16_t> xperm(x.size()), want2(x.size());
  for (int b = 0; b < B; ++b)
    for (int h = 0; h < H; ++h)
      for (int t = 0; t < T; ++t)
        memcpy(&xperm[(((size_t)b * H + h) * T + t) * D], &x[(((size_t)b * T + t) * H + h) * D], D * 2);
  omx_oracle_rope(xperm.data(), want2.data(), OMX_BFLOAT16, B, H, T, D, D, 0, 1, 1e6f, 1.0f, offset, nullptr);
  EXPECT(download(o2) == want2, "rope on a transposed view differs");
  // base xor freqs (fast.rs:25-29)
  bool threw = false;
  try {
    omx::fast::rope(xa, D, false, std::nullopt, 1.0f, 0);
  } catch (const omx::Exception& e) {
    threw = std::string(e.what()).find("Only one of base or freqs") != std::string::npos;
  }
  EXPECT(threw, "rope without base and freqs must raise the reference's message");
}

static void test_rms_norm_bit_exact() {
  printf("test_rms_norm_bit_exact\n");
  const int rows = 37, D = 128;
  auto x = randn_bf16((size_t)rows * D, 2), w = randn_bf16(D, 3);
  Array xa = Array::from_host(x.data(), {rows, D}, Dtype::Bfloat16), wa = Array::from_host(w.data(), {D}, Dtype::Bfloat16);
  Array out = omx::nn::RmsNorm{wa, 1e-6f}.forward(xa);
  std::vector<uint16_t> want(x.size());
  omx_oracle_rms_norm(x.data(), w.data(), want.data(), OMX_BFLOAT16, rows, D, 1e-6f);
  EXPECT(download(out) == want, "rms_norm output differs from the oracle bit pattern");
}

static void test_kv_cache_appendix_a() {
  printf("test_kv_cache_appendix_a\n");
  const int B = 1, H = 2, D = 64;
  auto cap = [](const omx::KVCache& c) { return c.state().first.shape()[2]; };
  auto mk = [&](int n, unsigned seed) {
    auto h = randn_bf16((size_t)B * H * n * D, seed);
    return std::make_pair(Array::from_host(h.data(), {B, H, n, D}, Dtype::Bfloat16), h);
  };
  {  // A1..A3: 5, then 251 single tokens, then one more (cache.rs:141-174)
    omx::KVCache c;
    auto kv = mk(5, 10);
    auto f = c.update_and_fetch(kv.first, kv.first);
    EXPECT(c.offset() == 5 && cap(c) == 256 && f.first.shape()[2] == 5, "A1: (%d, %lld)", c.offset(), (long long)cap(c));
    for (int i = 0; i < 251; ++i) {
      auto one = mk(1, 100 + i);
      c.update_and_fetch(one.first, one.first);
    }
    EXPECT(c.offset() == 256 && cap(c) == 256, "A2: (%d, %lld)", c.offset(), (long long)cap(c));
    auto one = mk(1, 999);
    auto f2 = c.update_and_fetch(one.first, one.first);
    EXPECT(c.offset() == 257 && cap(c) == 512, "A3: (%d, %lld)", c.offset(), (long long)cap(c));
    EXPECT(f2.first.strides()[1] >= 257 * D && f2.first.strides()[2] == D, "fetched keys are a strided view of the buffer");
    auto rows = download(f2.first);  // first 5 rows of head 0 are the first append
    EXPECT(memcmp(rows.data(), kv.second.data(), 5 * D * 2) == 0, "appended rows must be bit copies");
  }
  {  // A4, A5: 300 + 300 -> cap 812 (old buffer trimmed to 300, + 512)
    omx::KVCache c;
    auto a = mk(300, 20);
    c.update_and_fetch(a.first, a.first);
    EXPECT(c.offset() == 300 && cap(c) == 512, "A4");
    c.update_and_fetch(a.first, a.first);
    EXPECT(c.offset() == 600 && cap(c) == 812, "A5: cap %lld", (long long)cap(c));
    c.reset();  // A6: reset only rewinds the offset
    EXPECT(c.offset() == 0 && cap(c) == 812 && !c.max_size().has_value(), "reset");
  }
  {  // ConcatKeyValueCache: offset = shape[-2] of the concatenation (cache.rs:66-84)
    omx::ConcatKeyValueCache c;
    auto a = mk(7, 30), b = mk(3, 31);
    c.update_and_fetch(a.first, a.first);
    auto f = c.update_and_fetch(b.first, b.first);
    EXPECT(c.offset() == 10 && f.first.shape()[2] == 10, "concat cache offset");
  }
}

static float max_abs_diff(const std::vector<uint16_t>& a, const std::vector<uint16_t>& b) {
  float m = 0.f;
  for (size_t i = 0; i < a.size(); ++i) m = std::max(m, std::fabs(bf2f(a[i]) - bf2f(b[i])));
  return m;
}

// float64 causal attention [B,H,L,D] (same q/k/v heads): the yardstick for the reference chain's own rounding noise
static std::vector<double> sdpa_causal_f64(const std::vector<uint16_t>& q, const std::vector<uint16_t>& k,
                                           const std::vector<uint16_t>& v, int B, int H, int L, int D, float scale) {
  std::vector<double> out((size_t)B * H * L * D, 0.0), p(L);
  for (int bh = 0; bh < B * H; ++bh) {
    const size_t base = (size_t)bh * L * D;
    for (int i = 0; i < L; ++i) {
      double mx = -1e300, sum = 0.0;
      for (int j = 0; j <= i; ++j) {
        double s = 0.0;
        for (int d = 0; d < D; ++d) s += (double)bf2f(q[base + (size_t)i * D + d]) * (double)bf2f(k[base + (size_t)j * D + d]);
        p[j] = s * (double)scale;
        mx = std::max(mx, p[j]);
      }
      for (int j = 0; j <= i; ++j) sum += (p[j] = std::exp(p[j] - mx));
      for (int j = 0; j <= i; ++j) {
        const double w = p[j] / sum;
        for (int d = 0; d < D; ++d) out[base + (size_t)i * D + d] += w * (double)bf2f(v[base + (size_t)j * D + d]);
      }
    }
  }
  return out;
}

static void test_sdpa_shapes_like_the_reference_test() {
  printf("test_sdpa_shapes_like_the_reference_test\n");
  // mlx-rs/src/fast.rs:301-331 checks shape and dtype for B2, H24, L in {63, 129, 400}, D64; here values too
  for (int L : {63, 129, 400}) {
    const int B = 2, H = 24, D = 64;
    auto q = randn_bf16((size_t)B * H * L * D, 40 + L), k = randn_bf16(q.size(), 41 + L), v = randn_bf16(q.size(), 42 + L);
    Array qa = Array::from_host(q.data(), {B, H, L, D}, Dtype::Bfloat16);
    Array ka = Array::from_host(k.data(), {B, H, L, D}, Dtype::Bfloat16);
    Array va = Array::from_host(v.data(), {B, H, L, D}, Dtype::Bfloat16);
    const float scale = 1.0f / std::sqrt((float)D);
    Array out = omx::fast::scaled_dot_product_attention(qa, ka, va, scale, omx::fast::ScaledDotProductAttentionMask::causal());
    EXPECT(out.shape() == qa.shape() && out.dtype() == Dtype::Bfloat16, "sdpa shape/dtype");
    std::vector<uint16_t> want(q.size());
    omx_oracle_sdpa(q.data(), k.data(), v.data(), want.data(), OMX_BFLOAT16, B, H, H, L, L, D, D, scale, 1, nullptr, 0,
                    nullptr, 0);
    // 2e-2 max-abs against the reference chain (north_star).  That chain rounds q*scale, the scores and the
    // probabilities to bf16, so on the short causal rows of these 48 heads (1-3 keys: nothing averages out) it
    // sits up to 0.020 from the float64 answer itself, while the kernel (f32 scores, one final rounding) stays
    // within 0.009 (scripts/exp_fmha_err.py).  An element may therefore exceed 2e-2 against the chain ONLY where
    // the kernel is at least as close to the float64 answer as the chain is; everywhere else the bar is 2e-2.
    const std::vector<uint16_t> got = download(out);
    const std::vector<double> exact = sdpa_causal_f64(q, k, v, B, H, L, D, scale);
    float err = 0.f;
    size_t excused = 0;
    for (size_t i = 0; i < got.size(); ++i) {
      const float r = bf2f(want[i]), g = bf2f(got[i]), d = std::fabs(g - r);
      if (d > 2e-2f && std::fabs((double)g - exact[i]) <= std::fabs((double)r - exact[i])) {
        ++excused;
        continue;
      }
      err = std::max(err, d);
    }
    EXPECT(err <= 2e-2f, "sdpa L=%d max-abs err %g > 2e-2", L, err);
    EXPECT(excused <= got.size() / 100000 + 2, "sdpa L=%d: %zu elements beyond 2e-2 of the chain (closer to float64)", L,
           excused);
  }
  bool threw = false;
  try {  // n_heads % n_kv_heads
    Array q = Array::empty({1, 6, 4, 64}, Dtype::Bfloat16), k = Array::empty({1, 4, 4, 64}, Dtype::Bfloat16);
    omx::fast::scaled_dot_product_attention(q, k, k, 1.f);
  } catch (const omx::Exception& e) {
    threw = std::string(e.what()).find("n_heads must be a multiple of n_kv_heads") != std::string::npos;
  }
  EXPECT(threw, "GQA divisibility error message");
}

static void test_attention_forward_prefill_then_decode() {
  printf("test_attention_forward_prefill_then_decode\n");
  // Attention::forward (qwen3-mlx/src/model.rs:161-215): prefill 300 tokens (causal), then one decode step,
  // fused composites vs the op-by-op chain through the same C++ API, and vs the oracle.
  const int B = 2, Hq = 8, Hkv = 2, D = 128, L = 300;
  const float scale = 1.0f / std::sqrt((float)D), eps = 1e-6f;
  auto qh = randn_bf16((size_t)B * L * Hq * D, 50), kh = randn_bf16((size_t)B * L * Hkv * D, 51),
       vh = randn_bf16((size_t)B * L * Hkv * D, 52);
  auto qw = randn_bf16(D, 53), kw = randn_bf16(D, 54);
  // projections are [B, L, H, D]; the crates view them [B, H, L, D]
  Array q = Array::from_host(qh.data(), {B, L, Hq, D}, Dtype::Bfloat16).transpose_axes({0, 2, 1, 3});
  Array k = Array::from_host(kh.data(), {B, L, Hkv, D}, Dtype::Bfloat16).transpose_axes({0, 2, 1, 3});
  Array v = Array::from_host(vh.data(), {B, L, Hkv, D}, Dtype::Bfloat16).transpose_axes({0, 2, 1, 3});
  omx::nn::Rope rope = omx::utils::initialize_rope(D, 1e6f, false);
  omx::nn::RmsNorm q_norm{Array::from_host(qw.data(), {D}, Dtype::Bfloat16), eps};
  omx::nn::RmsNorm k_norm{Array::from_host(kw.data(), {D}, Dtype::Bfloat16), eps};

  // (a) op by op, exactly the reference's sequence
  omx::KVCache c1;
  Array qn = q_norm.forward(q), kn = k_norm.forward(k);
  Array qr = rope.forward(qn, c1.offset()), kr = rope.forward(kn, c1.offset());
  auto kv = c1.update_and_fetch(kr, v);
  Array o1 = omx::utils::scaled_dot_product_attention(qr, kv.first, kv.second, &c1, scale, omx::utils::SdpaMask::causal());
  // (b) composite
  omx::KVCache c2;
  Array o2 = omx::utils::attention_prefill_fused(q, k, v, c2, &rope, scale, omx::utils::SdpaMask::causal(), &q_norm, &k_norm);
  EXPECT(c1.offset() == L && c2.offset() == L, "offsets after prefill");
  EXPECT(download(c1.state().first) == download(c2.state().first), "KV keys: composite vs op-by-op must be bit-identical");
  EXPECT(download(c1.state().second) == download(c2.state().second), "KV values: composite vs op-by-op");
  EXPECT(max_abs_diff(download(o1), download(o2)) == 0.f, "prefill outputs: same kernels, same bits");

  // decode step
  auto q1h = randn_bf16((size_t)B * Hq * D, 60), k1h = randn_bf16((size_t)B * Hkv * D, 61), v1h = randn_bf16((size_t)B * Hkv * D, 62);
  Array q1 = Array::from_host(q1h.data(), {B, Hq, 1, D}, Dtype::Bfloat16);
  Array k1 = Array::from_host(k1h.data(), {B, Hkv, 1, D}, Dtype::Bfloat16);
  Array v1 = Array::from_host(v1h.data(), {B, Hkv, 1, D}, Dtype::Bfloat16);
  Array q1r = rope.forward(q_norm.forward(q1), c1.offset()), k1r = rope.forward(k_norm.forward(k1), c1.offset());
  auto kv1 = c1.update_and_fetch(k1r, v1);
  Array d1 = omx::utils::scaled_dot_product_attention(q1r, kv1.first, kv1.second, &c1, scale);
  Array d2 = omx::utils::attention_decode_fused(q1, k1, v1, c2, &rope, scale, &q_norm, &k_norm);
  EXPECT(std::string(omx_last_kernel()) == "decode_hmma_tma", "fused decode kernel: %s", omx_last_kernel());
  EXPECT(c1.offset() == L + 1 && c2.offset() == L + 1, "offsets after decode");
  EXPECT(download(c1.state().first) == download(c2.state().first), "KV keys after the fused decode step");
  const float derr = max_abs_diff(download(d1), download(d2));
  EXPECT(derr <= 2e-2f, "decode: fused vs op-by-op max-abs %g", derr);

  // (c) the decode step against the oracle chain on the host
  std::vector<uint16_t> qn1(q1h.size()), kn1(k1h.size()), qro(q1h.size()), kro(k1h.size());
  omx_oracle_rms_norm(q1h.data(), qw.data(), qn1.data(), OMX_BFLOAT16, B * Hq, D, eps);
  omx_oracle_rms_norm(k1h.data(), kw.data(), kn1.data(), OMX_BFLOAT16, B * Hkv, D, eps);
  omx_oracle_rope(qn1.data(), qro.data(), OMX_BFLOAT16, B, Hq, 1, D, D, 0, 1, 1e6f, 1.0f, L, nullptr);
  omx_oracle_rope(kn1.data(), kro.data(), OMX_BFLOAT16, B, Hkv, 1, D, D, 0, 1, 1e6f, 1.0f, L, nullptr);
  auto Kall = download(kv1.first), Vall = download(kv1.second);  // [B, Hkv, L+1, D] contiguous copies
  bool krow_ok = true;
  for (int b = 0; b < B && krow_ok; ++b)
    for (int h = 0; h < Hkv; ++h)
      krow_ok = krow_ok && memcmp(&Kall[(((size_t)b * Hkv + h) * (L + 1) + L) * D], &kro[((size_t)b * Hkv + h) * D], D * 2) == 0;
  EXPECT(krow_ok, "appended key row != oracle rope(rms_norm(k_new)) bit pattern");
  std::vector<uint16_t> want(q1h.size());
  omx_oracle_sdpa(qro.data(), Kall.data(), Vall.data(), want.data(), OMX_BFLOAT16, B, Hq, Hkv, 1, L + 1, D, D, scale, 0,
                  nullptr, 0, nullptr, 0);
  const float oerr = max_abs_diff(download(d2), want);
  EXPECT(oerr <= 2e-2f, "fused decode vs oracle max-abs %g", oerr);
}

// The decode loop as ONE CUDA graph from a compiled host: two layers, the position lives in
// device memory, one capture, five replays; every replay must reproduce the eager fused step bit for bit.
#define STAGE(msg) do { fprintf(stderr, "[graph loop test] %s\n", msg); fflush(stderr); } while (0)
static void test_decode_loop_cuda_graph() {
  const int B = 1, Hq = 16, Hkv = 8, D = 128, S0 = 250, LAYERS = 2, STEPS = 8;  // crosses the 256-row growth
  const float scale = 1.0f / std::sqrt((float)D), eps = 1e-6f;
  cudaStream_t st;
  cudaStreamCreate(&st);
  omx::Stream S{st};
  omx::nn::Rope rope = omx::utils::initialize_rope(D, 1e6f, false);
  auto qw = randn_bf16(D, 70), kw = randn_bf16(D, 71);
  omx::nn::RmsNorm q_norm{Array::from_host(qw.data(), {D}, Dtype::Bfloat16, S), eps};
  omx::nn::RmsNorm k_norm{Array::from_host(kw.data(), {D}, Dtype::Bfloat16, S), eps};
  std::vector<omx::KVCache> eager, graphed;
  std::vector<Array> q, k, v, out;
  for (int l = 0; l < LAYERS; ++l) {
    auto kh = randn_bf16((size_t)B * Hkv * S0 * D, 80 + l), vh = randn_bf16((size_t)B * Hkv * S0 * D, 90 + l);
    Array k0 = Array::from_host(kh.data(), {B, Hkv, S0, D}, Dtype::Bfloat16, S);
    Array v0 = Array::from_host(vh.data(), {B, Hkv, S0, D}, Dtype::Bfloat16, S);
    eager.emplace_back(256, S);
    graphed.emplace_back(256, S);
    eager.back().update_and_fetch(k0, v0);
    graphed.back().update_and_fetch(k0, v0);
    graphed.back().prepare_graph(S0 + STEPS, Hq);
    q.push_back(Array::empty({B, Hq, 1, D}, Dtype::Bfloat16));
    k.push_back(Array::empty({B, Hkv, 1, D}, Dtype::Bfloat16));
    v.push_back(Array::empty({B, Hkv, 1, D}, Dtype::Bfloat16));
    out.push_back(Array::empty({B, Hq, 1, D}, Dtype::Bfloat16));
  }
  int32_t* pos = nullptr;
  cudaMalloc(&pos, sizeof(int32_t));
  const int32_t p0 = S0;
  cudaMemcpyAsync(pos, &p0, sizeof(p0), cudaMemcpyHostToDevice, st);
  auto step = [&] {
    for (int l = 0; l < LAYERS; ++l)
      omx::utils::attention_decode_fused_dynamic(out[l], q[l], k[l], v[l], graphed[l], &rope, scale, pos, &q_norm, &k_norm, S);
    omx::utils::device_counter_add(pos, 1, S);
  };
  auto upload = [&](int t) {
    for (int l = 0; l < LAYERS; ++l) {
      auto a = randn_bf16((size_t)B * Hq * D, 1000 + 10 * t + l), b = randn_bf16((size_t)B * Hkv * D, 2000 + 10 * t + l),
           c = randn_bf16((size_t)B * Hkv * D, 3000 + 10 * t + l);
      cudaMemcpyAsync(q[l].data(), a.data(), a.size() * 2, cudaMemcpyHostToDevice, st);
      cudaMemcpyAsync(k[l].data(), b.data(), b.size() * 2, cudaMemcpyHostToDevice, st);
      cudaMemcpyAsync(v[l].data(), c.data(), c.size() * 2, cudaMemcpyHostToDevice, st);
      cudaStreamSynchronize(st);  // the host vectors die at the end of the iteration
    }
  };
  STAGE("setup done");
  upload(0);
  STAGE("uploaded");
  step();                                  // eager warm-up: builds the rope table ...
  omx::utils::device_counter_add(pos, -1, S);  // ... and is rolled back (row S0 is rewritten by the first replay)
  STAGE("warm-up done");
  cudaGraph_t graph;
  cudaGraphExec_t exec;
  EXPECT(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess, "begin capture");
  step();
  STAGE("captured");
  EXPECT(cudaStreamEndCapture(st, &graph) == cudaSuccess, "end capture");
  EXPECT(cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess, "instantiate");
  STAGE("instantiated");
  size_t n_nodes = 0;
  cudaGraphGetNodes(graph, nullptr, &n_nodes);
  EXPECT(n_nodes == (size_t)LAYERS + 1, "captured graph has %zu nodes, expected one launch per layer + the counter", n_nodes);
  bool same = true;
  for (int t = 0; t < STEPS; ++t) {
    upload(t);
    EXPECT(cudaGraphLaunch(exec, st) == cudaSuccess, "graph launch");
    for (int l = 0; l < LAYERS; ++l) {
      graphed[l].advance(1);
      Array want = omx::utils::attention_decode_fused(q[l], k[l], v[l], eager[l], &rope, scale, &q_norm, &k_norm, S);
      cudaStreamSynchronize(st);
      same = same && download(want) == download(out[l]);
    }
  }
  STAGE("replayed");
  EXPECT(same, "graph replays must reproduce the eager fused step bit for bit");
  int32_t pend = 0;
  cudaMemcpy(&pend, pos, sizeof(pend), cudaMemcpyDeviceToHost);
  EXPECT(pend == S0 + STEPS, "device position %d after %d replays", pend, STEPS);
  for (int l = 0; l < LAYERS; ++l) {
    EXPECT(graphed[l].offset() == S0 + STEPS && eager[l].offset() == S0 + STEPS, "offsets after the loop");
    EXPECT(download(graphed[l].state().first) == download(eager[l].state().first), "KV keys after graph replays (layer %d)", l);
    EXPECT(download(graphed[l].state().second) == download(eager[l].state().second), "KV values after graph replays (layer %d)", l);
  }
  cudaGraphExecDestroy(exec);
  cudaGraphDestroy(graph);
  cudaFree(pos);
  // the caches free their buffers stream-ordered on the stream they were used on: drop them before it
  eager.clear();
  graphed.clear();
  cudaStreamSynchronize(st);
  cudaStreamDestroy(st);
}

static void test_paged_cache_equals_contiguous_cache() {
  printf("test_paged_cache_equals_contiguous_cache\n");
  // the decode loop of Attention::forward over a PagedKVCache and over the reference-shaped KVCache: same kernels,
  // same arithmetic -> outputs and cache rows bit-identical, across a page boundary (120 -> 136 rows)
  const int B = 2, Hq = 8, Hkv = 2, D = 128, S0 = 120, steps = 16;
  const float scale = 1.0f / std::sqrt((float)D);
  auto kh = randn_bf16((size_t)B * Hkv * S0 * D, 70), vh = randn_bf16((size_t)B * Hkv * S0 * D, 71);
  Array k0 = Array::from_host(kh.data(), {B, Hkv, S0, D}, Dtype::Bfloat16);
  Array v0 = Array::from_host(vh.data(), {B, Hkv, S0, D}, Dtype::Bfloat16);
  omx::nn::Rope rope = omx::utils::initialize_rope(D, 1e6f, false);
  omx::KVCache cc;
  omx::PagedKVCache pc(B, Hkv, D, Dtype::Bfloat16, /*n_pages=*/8, /*max_pages_per_seq=*/4);
  cc.update_and_fetch(k0, v0);
  auto pf = pc.update_and_fetch(k0, v0);
  EXPECT(download(pf.first) == kh && download(pf.second) == vh, "materialised prefill rows == the appended rows");
  EXPECT(pc.offset() == S0 && pc.free_pages() == 8 - 2 * 2, "paged offsets / pages after prefill");
  for (int t = 0; t < steps; ++t) {
    auto qh = randn_bf16((size_t)B * Hq * D, 80 + t), k1h = randn_bf16((size_t)B * Hkv * D, 100 + t),
         v1h = randn_bf16((size_t)B * Hkv * D, 120 + t);
    Array q = Array::from_host(qh.data(), {B, Hq, 1, D}, Dtype::Bfloat16);
    Array k1 = Array::from_host(k1h.data(), {B, Hkv, 1, D}, Dtype::Bfloat16);
    Array v1 = Array::from_host(v1h.data(), {B, Hkv, 1, D}, Dtype::Bfloat16);
    Array a = omx::utils::attention_decode_fused(q, k1, v1, cc, &rope, scale);
    Array b = omx::utils::attention_decode_fused_paged(q, k1, v1, pc, &rope, scale);
    EXPECT(download(a) == download(b), "step %d: paged and contiguous outputs must be bit-identical", t);
  }
  EXPECT(pc.offset() == S0 + steps && pc.free_pages() == 8 - 2 * 3, "a third page per sequence after the boundary");
  auto kv = pc.fetch();
  const auto ck = download(cc.state().first), pk = download(kv.first);
  // contiguous state is [B,Hkv,cap,D] with cap >= offset; compare the first offset rows of every (b, h)
  const int cap = (int)cc.state().first.shape()[2], n = S0 + steps;
  bool same = true;
  for (int bh = 0; bh < B * Hkv && same; ++bh)
    same = std::equal(pk.begin() + (size_t)bh * n * D, pk.begin() + (size_t)(bh + 1) * n * D, ck.begin() + (size_t)bh * cap * D);
  EXPECT(same, "paged rows == contiguous cache rows, bit for bit");
  pc.release(1);
  EXPECT(pc.lengths()[1] == -1 && pc.lengths()[0] == n && pc.free_pages() == 8 - 3, "release returns the pages");
}

// The head-sharded step with the data + flag exchange at world = 1 (the exchange degenerates to the local store
// and the step counter): same bits as the plain fused step, counter = number of steps.
static void test_head_sharded_ll_world1() {
  printf("test_head_sharded_ll_world1\n");
  const int B = 1, Hq = 4, Hkv = 1, D = 128, L = 3000;
  const float scale = 1.0f / std::sqrt((float)D);
  auto kh = randn_bf16((size_t)B * Hkv * L * D, 70), vh = randn_bf16((size_t)B * Hkv * L * D, 71);
  Array k = Array::from_host(kh.data(), {B, Hkv, L, D}, Dtype::Bfloat16);
  Array v = Array::from_host(vh.data(), {B, Hkv, L, D}, Dtype::Bfloat16);
  omx::nn::Rope rope = omx::utils::initialize_rope(D, 1e6f, false);
  omx::KVCache c1, c2;
  c1.update_and_fetch(k, v);
  c2.update_and_fetch(k, v);
  const size_t sb = omx_ll_staging_bytes(1, B, Hq, D, OMX_BFLOAT16);
  EXPECT(sb == (size_t)2 * B * Hq * D * 2 / 4 * 8, "staging bytes %zu", sb);
  void* staging = nullptr;
  uint32_t* seq = nullptr;
  cudaMalloc(&staging, sb);
  cudaMemset(staging, 0, sb);
  cudaMalloc(&seq, 4);
  cudaMemset(seq, 0, 4);
  omx_ll_group g{};
  g.world = 1;
  g.rank = 0;
  g.staging[0] = staging;
  g.seq = seq;
  Array out_full = Array::empty({B, Hq, 1, D}, Dtype::Bfloat16);
  for (int step = 1; step <= 3; ++step) {
    auto qh = randn_bf16((size_t)B * Hq * D, 72 + step), k1h = randn_bf16((size_t)B * Hkv * D, 80 + step),
         v1h = randn_bf16((size_t)B * Hkv * D, 90 + step);
    Array q1 = Array::from_host(qh.data(), {B, Hq, 1, D}, Dtype::Bfloat16);
    Array k1 = Array::from_host(k1h.data(), {B, Hkv, 1, D}, Dtype::Bfloat16);
    Array v1 = Array::from_host(v1h.data(), {B, Hkv, 1, D}, Dtype::Bfloat16);
    Array ref = omx::utils::attention_decode_fused(q1, k1, v1, c1, &rope, scale);
    omx::utils::attention_decode_fused_sharded(out_full, q1, k1, v1, c2, &rope, scale, g);
    EXPECT(download(ref) == download(out_full), "step %d: sharded (world 1) output differs from the plain fused step", step);
    uint32_t hs = 0;
    cudaMemcpy(&hs, seq, 4, cudaMemcpyDeviceToHost);
    EXPECT((int)hs == step, "step counter %u after %d steps", hs, step);
  }
  EXPECT(download(c1.state().first) == download(c2.state().first), "KV keys: sharded vs plain");
  cudaFree(staging);
  cudaFree(seq);
}

// --bench: the fused decode step launched EAGERLY from compiled code (no Python, no graph): what a Rust / C++ host
// pays per layer and token.  Caches are rotated so that every step reads HBM; the offset is rewound after each step.
static void bench_eager_decode(const char* name, int B, int Hq, int Hkv, int S, Dtype dt, int R) {
  const int D = 128;
  const size_t es = dt == Dtype::Float32 ? 4 : 2;
  const float scale = 1.0f / std::sqrt((float)D);
  std::vector<uint16_t> h16;
  std::vector<float> h32;
  auto host = [&](size_t n, unsigned seed) -> const void* {
    if (es == 2) {
      h16 = randn_bf16(n, seed);
      return h16.data();
    }
    std::mt19937 g(seed);
    std::normal_distribution<float> d(0.f, 1.f);
    h32.resize(n);
    for (auto& x : h32) x = d(g);
    return h32.data();
  };
  Array k = Array::from_host(host((size_t)B * Hkv * (S - 1) * D, 1), {B, Hkv, S - 1, D}, dt);
  Array v = Array::from_host(host((size_t)B * Hkv * (S - 1) * D, 2), {B, Hkv, S - 1, D}, dt);
  Array q1 = Array::from_host(host((size_t)B * Hq * D, 3), {B, Hq, 1, D}, dt);
  Array k1 = Array::from_host(host((size_t)B * Hkv * D, 4), {B, Hkv, 1, D}, dt);
  Array v1 = Array::from_host(host((size_t)B * Hkv * D, 5), {B, Hkv, 1, D}, dt);
  Array out = Array::empty({B, Hq, 1, D}, dt);
  std::vector<omx::KVCache> caches;
  for (int r = 0; r < R; ++r) {
    caches.emplace_back();
    omx_kv_cache_reserve(caches.back().raw(), S + 256);
    caches.back().update_and_fetch(k, v);
  }
  omx_optional_float base{1e6f, true};
  auto step = [&](int i) {
    omx::KVCache& c = caches[i % R];
    omx::check(omx_attn_decode_fused(out.desc(), q1.desc(), k1.desc(), v1.desc(), c.raw(), D, false, base, 1.0f, nullptr,
                                     scale, nullptr, nullptr, nullptr));
    int t = 0;
    omx_kv_cache_trim(c.raw(), 1, &t);
  };
  for (int i = 0; i < 3 * R; ++i) step(i);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int N = 50 * R;
  cudaEventRecord(e0);
  for (int i = 0; i < N; ++i) step(i);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  printf("{\"host\": \"c++ eager\", \"shape\": \"%s\", \"us_per_step\": %.2f, \"kernel\": \"%s\", \"caches\": %d}\n", name,
         ms * 1e3 / N, omx_last_kernel(), R);
}

int main(int argc, char** argv) {
  int sm = 0;
  if (omx_device_check(&sm) != 0) {
    printf("no sm_100a device: %s\n", omx_last_error());
    return 2;
  }
  if (argc > 1 && std::string(argv[1]) == "--bench") {
    try {
      bench_eager_decode("c1 fp32 16/8 ctx2048", 1, 16, 8, 2048, Dtype::Float32, 16);
      bench_eager_decode("0.6b bf16 16/8 ctx2048", 1, 16, 8, 2048, Dtype::Bfloat16, 16);
      bench_eager_decode("8b bf16 32/8 ctx8192", 1, 32, 8, 8192, Dtype::Bfloat16, 16);
      bench_eager_decode("c5 bf16 32/8 ctx32768", 1, 32, 8, 32768, Dtype::Bfloat16, 4);
      bench_eager_decode("c2 bf16 32/8 ctx8192 B8", 8, 32, 8, 8192, Dtype::Bfloat16, 1);
    } catch (const std::exception& e) {
      printf("EXCEPTION: %s\n", e.what());
      return 1;
    }
    return 0;
  }
  try {
    test_rope_bit_exact();
    test_rms_norm_bit_exact();
    test_kv_cache_appendix_a();
    test_sdpa_shapes_like_the_reference_test();
    test_attention_forward_prefill_then_decode();
    test_decode_loop_cuda_graph();
    test_paged_cache_equals_contiguous_cache();
    test_head_sharded_ll_world1();
  } catch (const std::exception& e) {
    printf("EXCEPTION: %s\n", e.what());
    return 1;
  }
  cudaDeviceSynchronize();
  printf(g_failed ? "%d check(s) FAILED\n" : "all host-layer checks passed\n", g_failed);
  return g_failed ? 1 : 0;
}
