// Legacy tensor path microbenchmark (B200): what does `mma.sync.m16n8k16` (bf16, f32 accumulate; SASS HMMA.16816.F32.BF16)
// sustain per SM, alone and fed by `ldmatrix.x4` from shared memory?  W warps per SM (one CTA per SM), each runs
// ITER x 16 MMAs on 8 independent accumulators; clock64 around the loop of warp 0 of every CTA.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hmma hmma.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}

// MODE 0: MMAs only (operands in registers).  MODE 1: per 4 MMAs one A ldmatrix + two B ldmatrix (the 16 x 32 S tile
// of sdpa_mma).  MODE 2: per 2 MMAs one ldmatrix (its PV step).  MODE 3: ldmatrix only.
template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
  extern __shared__ __align__(16) unsigned char sm[];
  for (int i = threadIdx.x; i < 16384; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0x3f803f80u;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm) + (warp & 3) * 4096 + (lane & 15) * 144 + (lane >> 4) * 16;
  float acc[8][4] = {};
  uint32_t a[4] = {0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u}, b[4] = {0x3f803f80u, 0x3c003c00u, 0x3f803f80u, 0x3c003c00u};
  uint32_t x = 0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (MODE == 1) {
        uint32_t b2[4];
        ldsm4(base + u * 32, a);
        ldsm4(base + 2304 + u * 32, b);
        ldsm4(base + 4608 + u * 32, b2);
        mma(acc[0], a, b[0], b[1]); mma(acc[1], a, b[2], b[3]); mma(acc[2], a, b2[0], b2[1]); mma(acc[3], a, b2[2], b2[3]);
      } else if (MODE == 2) {
        uint32_t b2[4];
        ldsm4(base + u * 32, b);
        ldsm4(base + 2304 + u * 32, b2);
        mma(acc[0 + (u & 1) * 4], a, b[0], b[1]); mma(acc[1 + (u & 1) * 4], a, b[2], b[3]);
        mma(acc[2 + (u & 1) * 4], a, b2[0], b2[1]); mma(acc[3 + (u & 1) * 4], a, b2[2], b2[3]);
      } else if (MODE == 3) {
        uint32_t r0[4], r1[4], r2[4];
        ldsm4(base + u * 32, r0); ldsm4(base + 2304 + u * 32, r1); ldsm4(base + 4608 + u * 32, r2);
        x ^= r0[0] ^ r0[3] ^ r1[1] ^ r1[2] ^ r2[0] ^ r2[3];
      } else {
        mma(acc[2 * (u & 3)], a, b[0], b[1]); mma(acc[2 * (u & 3) + 1], a, b[2], b[3]);
        mma(acc[(2 * u + 4) & 7], a, b[0], b[1]); mma(acc[(2 * u + 5) & 7], a, b[2], b[3]);
      }
    }
  }
  const long long t1 = clock64();
  float s = __uint_as_float(x);
  for (int i = 0; i < 8; ++i) s += acc[i][0] + acc[i][1] + acc[i][2] + acc[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, float* out, long long* cyc) {
  const int sms = 148, iters = 4096;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  for (int warps : {4, 8, 16}) {
    k<MODE><<<sms, warps * 32, 65536>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<sms, warps * 32, 65536>>>(out, cyc, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < sms; ++i) c += h[i]; c /= sms;
    const double mmas = MODE == 3 ? 0 : (double)iters * 16 * warps;       // per SM
    const double ldsm = MODE == 0 ? 0 : (double)iters * 4 * (MODE == 2 ? 2 : 3) * warps;
    printf("%-28s warps/SM %2d: %8.0f cycles  %6.2f cyc/MMA/SM  %7.1f FLOP/clk/SM  %6.1f TFLOP/s (events)  ldmatrix.x4 %5.2f cyc each/SM (%5.1f B/clk)\n",
           name, warps, c, mmas ? c / mmas : 0.0, mmas * 4096 / c, mmas * 4096 * sms / (ms * 1e-3) / 1e12,
           ldsm ? c / ldsm : 0.0, ldsm ? ldsm * 512 / c : 0.0);
  }
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  run<0>("mma only", out, cyc);
  run<1>("1 A + 2 B ldmatrix / 4 mma", out, cyc);
  run<2>("1 ldmatrix / 2 mma", out, cyc);
  run<3>("ldmatrix only", out, cyc);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
