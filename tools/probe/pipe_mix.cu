// Micro-benchmark (development tool, not product): issue rates of the packed min/max candidates for the FAST arc network on
// sm_100a, alone and interleaved, to see which ones share a pipe.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 pipe_mix.cu
#include <cstdio>
#include <cuda_runtime.h>
#define A_VIMNMX3(x, y, z) asm volatile("{.reg .b32 t; vimnmx3... }")
template <int MODE> __device__ __forceinline__ void step(unsigned (&r)[8], unsigned k) {
#pragma unroll
  for (int i = 0; i < 8; i++) {
    unsigned a = r[i], b = r[(i + 1) & 7], c = r[(i + 3) & 7];
    if (MODE == 0) r[i] = __vimax3_s16x2(a, b, c);                                           // VIMNMX3.S16x2
    if (MODE == 1) asm volatile("max.f16x2 %0, %1, %2;" : "=r"(r[i]) : "r"(a), "r"(b));        // HMNMX2
    if (MODE == 2) asm volatile("fma.rn.relu.f16x2 %0, %1, %2, %3;" : "=r"(r[i]) : "r"(a), "r"(b), "r"(c));   // HFMA2.RELU
    if (MODE == 3) r[i] = __byte_perm(a, b, 0x4140 + (k & 1));                                   // PRMT
    if (MODE == 4) r[i] = a * k + c;                                                             // IMAD
    if (MODE == 5) r[i] = __shfl_up_sync(0xffffffffu, a, 1);                                     // SHFL
    if (MODE == 6) r[i] = __vmaxs2(a, b);                                                        // VIMNMX.S16x2
    if (MODE == 7) r[i] = __viaddmax_s16x2_relu(a, b, c);                                        // VIADDMNMX
    if (MODE == 8) asm volatile("add.f16x2 %0, %1, %2;" : "=r"(r[i]) : "r"(a), "r"(b));         // HADD2
    if (MODE == 9) r[i] = __vadd2(a, b);                                                         // VIADD.16x2
    if (MODE == 10) asm volatile("max.f32 %0, %1, %2, %3;" : "=r"(r[i]) : "r"(a), "r"(b), "r"(c));  // FMNMX3
    if (MODE == 11) r[i] = (a & b) ^ c;                                                          // LOP3
  }
}
template <int M1, int M2, int N1, int N2> __global__ void __launch_bounds__(256) k_mix(unsigned *out, int iters, unsigned k) {
  unsigned r[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { r[i] = threadIdx.x * 2654435761u + i * k; q[i] = r[i] ^ 0x3c003c00u; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
#pragma unroll
      for (int a = 0; a < N1; a++) step<M1>(r, k);
#pragma unroll
      for (int a = 0; a < N2; a++) step<M2>(q, k);
    }
  }
  unsigned s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s ^= r[i] ^ q[i];
  if (s == 0x12345u) out[0] = s;
}
template <int M1, int M2, int N1, int N2> void run(const char *name) {
  unsigned *d; cudaMalloc(&d, 4);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int iters = 2000, blocks = sms * 8;
  k_mix<M1, M2, N1, N2><<<blocks, 256>>>(d, 10, 3);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k_mix<M1, M2, N1, N2><<<blocks, 256>>>(d, iters, 3);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double warpInst = (double)blocks * 8 * iters * 4 * 8 * (N1 + N2);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("%-44s %8.3f ms  %7.2f warp-inst/clk/SM (at %d MHz)\n", name, ms, warpInst / (ms * 1e-3) / (clk * 1e3) / sms, clk / 1000);
  cudaFree(d);
}
int main() {
  run<0, 0, 1, 0>("VIMNMX3.S16x2");
  run<6, 6, 1, 0>("VIMNMX.S16x2");
  run<7, 7, 1, 0>("VIADDMNMX.S16x2.RELU");
  run<1, 1, 1, 0>("HMNMX2");
  run<2, 2, 1, 0>("HFMA2.RELU");
  run<8, 8, 1, 0>("HADD2");
  run<3, 3, 1, 0>("PRMT");
  run<4, 4, 1, 0>("IMAD");
  run<5, 5, 1, 0>("SHFL.UP");
  run<9, 9, 1, 0>("VIADD.16x2");
  run<10, 10, 1, 0>("FMNMX3");
  run<11, 11, 1, 0>("LOP3");
  run<0, 1, 1, 1>("VIMNMX3 + HMNMX2 1:1");
  run<0, 2, 1, 1>("VIMNMX3 + HFMA2.RELU 1:1");
  run<0, 4, 1, 1>("VIMNMX3 + IMAD 1:1");
  run<0, 3, 1, 1>("VIMNMX3 + PRMT 1:1");
  run<0, 5, 1, 1>("VIMNMX3 + SHFL 1:1");
  run<0, 5, 4, 1>("VIMNMX3 + SHFL 4:1");
  run<1, 2, 1, 1>("HMNMX2 + HFMA2.RELU 1:1");
  run<1, 4, 1, 1>("HMNMX2 + IMAD 1:1");
  run<0, 10, 1, 1>("VIMNMX3 + FMNMX3 1:1");
  run<0, 9, 1, 1>("VIMNMX3 + VIADD.16x2 1:1");
  return 0;
}
