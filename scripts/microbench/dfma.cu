// fp64 FMA issue rate on B200: K independent chains per thread, N warps per CTA (1 CTA).
#include <cstdio>
#include <cuda_runtime.h>
template <int K>
__global__ void k(double *out, long long *cyc, double a, double b) {
  double y[K];
#pragma unroll
  for (int i = 0; i < K; ++i) y[i] = a + i + threadIdx.x;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < 64; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < K; ++i) y[i] = fma(y[i], b, a);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < K; ++i) s += y[i];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
template <int K> void run(double *out, long long *cyc) {
  for (int warps : {1, 2, 4, 8, 16, 32}) {
    k<K><<<1, 32 * warps>>>(out, cyc, 1.0000001, 0.9999999);
    cudaDeviceSynchronize();
    k<K><<<1, 32 * warps>>>(out, cyc, 1.0000001, 0.9999999);
    cudaDeviceSynchronize();
    double per = cyc[0] / (64.0 * 4 * K);
    printf("K=%2d chains  warps=%2d : %.2f cycles per DFMA per warp  -> %.1f fp64 FMA lanes/clk/SM\n", K, warps, per, 32.0 * warps / per);
  }
}
int main() {
  double *out; long long *cyc;
  cudaMalloc(&out, 2048 * 8); cudaMallocManaged(&cyc, 64);
  run<1>(out, cyc); run<4>(out, cyc); run<8>(out, cyc); run<36>(out, cyc);
  return 0;
}
