#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
template<int MODE>
__global__ void k(unsigned* out, unsigned seed, int iters) {
  unsigned a0 = seed + threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, a4=a0*11,a5=a0*13,a6=a0*17,a7=a0*19;
  for (int i = 0; i < iters; ++i) {
#define STEP(a) \
    if (MODE == 0) { float f = __uint_as_float(a); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f)); a = __float_as_uint(f); } \
    else if (MODE == 1) { asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(a)); } \
    else if (MODE == 2) { asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(a)); } \
    else if (MODE == 3) { asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(a)); }
    STEP(a0) STEP(a1) STEP(a2) STEP(a3) STEP(a4) STEP(a5) STEP(a6) STEP(a7)
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3^a4^a5^a6^a7;
}
template<int MODE> void run(const char* name, int vals_per_op) {
  unsigned* out; cudaMalloc(&out, 148 * 4 * 1024 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int iters = 4096;
  k<MODE><<<148 * 2, 512>>>(out, 1, iters); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<MODE><<<148 * 2, 512>>>(out, 1, iters); cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double ops = 148.0 * 2 * 512 * iters * 8;
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("%s: %.3f ms, %.2f Gop/s (lane-ops), %.2f lane-ops/clk/SM @%d kHz nominal, values/clk/SM %.2f\n", name, ms, ops / ms * 1e-6, ops / (ms * 1e-3) / 148 / (clk * 1e3), clk, vals_per_op * ops / (ms * 1e-3) / 148 / (clk * 1e3));
}
int main() { run<0>("f32", 1); run<1>("f16x2", 2); run<2>("bf16x2.ftz", 2); run<3>("f16x2.ftz", 2); return 0; }
