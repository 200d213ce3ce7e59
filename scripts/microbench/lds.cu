// Shared-memory LDS.128 cost under different address patterns (B200), 16 warps per CTA.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double *out, long long *cyc, int mode) {
  extern __shared__ double2 sm[];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = make_double2(i, 2 * i);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int base;
  if (mode == 0) base = warp * 64;                          // all lanes same address (broadcast)
  else if (mode == 1) base = warp * 64 + (lane / 6) * 18;   // 5-6 distinct 288-byte blocks per warp
  else if (mode == 2) base = warp * 64 + lane * 18;         // 32 distinct blocks, stride 288 B
  else base = warp * 64 + lane;                             // 32 consecutive 16-byte words
  double acc = 0;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < 64; ++it) {
#pragma unroll
    for (int u = 0; u < 18; ++u) { const double2 v = sm[(base + u + it) & 8191]; acc += v.x + v.y; }
  }
  long long t1 = clock64();
  out[threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  double *out; long long *cyc;
  cudaMalloc(&out, 2048 * 8); cudaMallocManaged(&cyc, 64);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 16);
  const char *names[] = {"all lanes one address", "6-lane groups, 288 B apart", "32 lanes, 288 B stride", "32 lanes consecutive"};
  for (int warps : {1, 4, 16})
    for (int mode = 0; mode < 4; ++mode) {
      k<<<1, 32 * warps, 8192 * 16>>>(out, cyc, mode); cudaDeviceSynchronize();
      k<<<1, 32 * warps, 8192 * 16>>>(out, cyc, mode); cudaDeviceSynchronize();
      printf("warps=%2d %-28s: %.2f cycles per LDS.128 per warp -> %.2f cycles of the SM pipe per warp-instruction\n", warps, names[mode],
             cyc[0] / (64.0 * 18), cyc[0] / (64.0 * 18) / warps);
    }
  return 0;
}
