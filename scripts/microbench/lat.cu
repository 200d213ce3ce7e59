// Dependent-issue latency micro-benchmarks on B200 (fp64 pipe, shuffles, shared memory).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double *out, long long *cyc, double a, double b) {
  __shared__ double sm[64];
  sm[threadIdx.x & 63] = a;
  __syncthreads();
  double x = a + threadIdx.x;
  long long t0, t1;
#define CLK(t) do { asm volatile("" : "+d"(x)); t = clock64(); asm volatile("" : "+d"(x)); } while (0)
  // DFMA chain
  CLK(t0);
#pragma unroll
  for (int i = 0; i < 256; ++i) x = fma(x, b, a);
  CLK(t1);
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  // DMUL chain
  CLK(t0);
#pragma unroll
  for (int i = 0; i < 256; ++i) x = x * b;
  CLK(t1);
  if (threadIdx.x == 0) cyc[1] = t1 - t0;
  // shuffle chain (64-bit)
  CLK(t0);
#pragma unroll
  for (int i = 0; i < 256; ++i) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31);
  CLK(t1);
  if (threadIdx.x == 0) cyc[2] = t1 - t0;
  // rsqrt chain
  x = fabs(x) + 1.0;
  CLK(t0);
#pragma unroll
  for (int i = 0; i < 64; ++i) x = rsqrt(x) + 1.0;
  CLK(t1);
  if (threadIdx.x == 0) cyc[3] = t1 - t0;
  // LDS chain (pointer chase through values)
  int idx = threadIdx.x & 63;
  CLK(t0);
#pragma unroll
  for (int i = 0; i < 256; ++i) idx = ((int)sm[idx] + idx + 1) & 63;
  x += idx;
  CLK(t1);
  if (threadIdx.x == 0) cyc[4] = t1 - t0;
  // 8 independent DFMA chains (throughput per warp)
  CLK(t0);
  double y0 = x, y1 = x + 1, y2 = x + 2, y3 = x + 3, y4 = x + 4, y5 = x + 5, y6 = x + 6, y7 = x + 7;
  CLK(t0);
#pragma unroll
  for (int i = 0; i < 128; ++i) { y0 = fma(y0, b, a); y1 = fma(y1, b, a); y2 = fma(y2, b, a); y3 = fma(y3, b, a); y4 = fma(y4, b, a); y5 = fma(y5, b, a); y6 = fma(y6, b, a); y7 = fma(y7, b, a); }
  x += y0 + y1 + y2 + y3 + y4 + y5 + y6 + y7;
  CLK(t1);
  if (threadIdx.x == 0) cyc[5] = t1 - t0;
  // division chain
  CLK(t0);
#pragma unroll
  for (int i = 0; i < 64; ++i) x = a / (x + 1.0);
  CLK(t1);
  if (threadIdx.x == 0) cyc[6] = t1 - t0;
  out[threadIdx.x] = x + idx + y0 + y1 + y2 + y3 + y4 + y5 + y6 + y7;
}
int main() {
  double *out; long long *cyc;
  cudaMalloc(&out, 1024 * 8); cudaMallocManaged(&cyc, 64 * 8);
  for (int warps = 1; warps <= 16; warps *= 4) {
    k<<<1, 32 * warps>>>(out, cyc, 1.0000001, 0.9999999);
    cudaDeviceSynchronize();
    printf("warps=%2d  DFMA dep %.1f  DMUL dep %.1f  SHFL64 dep %.1f  rsqrt+add dep %.1f  LDS dep %.1f  DFMA x8 indep %.2f/instr  div+add dep %.1f\n", warps,
           cyc[0] / 256.0, cyc[1] / 256.0, cyc[2] / 256.0, cyc[3] / 64.0, cyc[4] / 256.0, cyc[5] / 1024.0, cyc[6] / 64.0);
  }
  return 0;
}
