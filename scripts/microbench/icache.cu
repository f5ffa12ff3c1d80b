// Instruction-fetch microbenchmark (B200): how fast does an SM run straight-line code it has never seen, does the
// instruction cache survive a kernel boundary, and how much code fits?  Each kernel is NB blocks of 1024
// independent-chain FFMAs (16 KB of SASS per block); one warp per SM.  Reported per launch, back-to-back launches:
//   pass1 = %clock64 cycles of the first walk through the code, pass2 = of a second walk in the same launch (warm).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o icache icache.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int SEED>
__device__ __forceinline__ void block1024(float& a, float& b, float& c, float& d) {
#pragma unroll
  for (int i = 0; i < 256; ++i) {
    a = fmaf(a, 1.0001f + SEED * 1e-6f, 0.5f + i * 1e-3f);
    b = fmaf(b, 0.9999f + SEED * 1e-6f, 0.25f + i * 1e-3f);
    c = fmaf(c, 1.0002f + SEED * 1e-6f, 0.125f + i * 1e-3f);
    d = fmaf(d, 0.9998f + SEED * 1e-6f, 0.0625f + i * 1e-3f);
  }
}
template <int NB, int I = 0>
struct Walk {
  __device__ static void run(float& a, float& b, float& c, float& d) {
    block1024<I>(a, b, c, d);
    if constexpr (I + 1 < NB) Walk<NB, I + 1>::run(a, b, c, d);
  }
};
template <int NB>
__global__ void k(float* out, long long* cyc, int passes) {
  float a = threadIdx.x, b = 1.f, c = 2.f, d = 3.f;
  long long t[4];
  t[0] = clock64();
  for (int p = 0; p < passes; ++p) {  // same code, walked `passes` times
    Walk<NB>::run(a, b, c, d);
    t[p + 1 < 3 ? p + 1 : 3] = clock64();
  }
  if (threadIdx.x == 0) {
    cyc[blockIdx.x * 2] = t[1] - t[0];
    cyc[blockIdx.x * 2 + 1] = t[2] - t[1];
  }
  out[blockIdx.x * 32 + threadIdx.x] = a + b + c + d;
}
template <int NB>
void run(float* out, long long* cyc) {
  const int sms = 148;
  for (int i = 0; i < 3; ++i) k<NB><<<sms, 32>>>(out, cyc, 2);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int L = 50;
  cudaEventRecord(e0);
  for (int i = 0; i < L; ++i) k<NB><<<sms, 32>>>(out, cyc, 1);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms1; cudaEventElapsedTime(&ms1, e0, e1);
  cudaEventRecord(e0);
  for (int i = 0; i < L; ++i) k<NB><<<sms, 32>>>(out, cyc, 2);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms2; cudaEventElapsedTime(&ms2, e0, e1);
  long long h[2 * 148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double p1 = 0, p2 = 0;
  for (int i = 0; i < sms; ++i) { p1 += h[2 * i]; p2 += h[2 * i + 1]; }
  printf("code %4d KB | launch (1 pass) %7.2f us | launch (2 passes) %7.2f us | pass1 %8.0f cyc (%.2f cyc/instr) | pass2 %8.0f cyc (%.2f cyc/instr)\n",
         NB * 16, ms1 * 1e3 / L, ms2 * 1e3 / L, p1 / sms, p1 / sms / (NB * 1024.0), p2 / sms, p2 / sms / (NB * 1024.0));
}
int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 32 * 4); cudaMalloc(&cyc, 148 * 2 * 8);
  run<1>(out, cyc); run<2>(out, cyc); run<4>(out, cyc); run<8>(out, cyc); run<16>(out, cyc); run<32>(out, cyc);
  return 0;
}
