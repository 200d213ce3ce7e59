// Step-level micro-benchmarks for k_tree_solve on B200: what one lane / one lane group pays for the pieces on the
// critical path of an elimination step (everything timed with clock64 inside one CTA of 512 threads).
#include <cstdio>
#include <cuda_runtime.h>
#include "../../ssvio_b200/csrc/ssba_block_inverse.cuh"
using namespace ssba;

__device__ __forceinline__ bool chol6(double *D) {
  double a[36];
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int k = 0; k <= i; ++k) a[6 * i + k] = D[6 * i + k];
  bool bad = false;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    double dj = a[7 * j];
    if (!(dj > 0.0)) { bad = true; dj = 1.0; }
    const double inv = rsqrt(dj);
    a[7 * j] = inv;
#pragma unroll
    for (int i = j + 1; i < 6; ++i) a[6 * i + j] *= inv;
#pragma unroll
    for (int i = j + 1; i < 6; ++i)
#pragma unroll
      for (int k = j + 1; k <= i; ++k) a[6 * i + k] -= a[6 * i + j] * a[6 * k + j];
  }
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int k = 0; k <= i; ++k) D[6 * i + k] = a[6 * i + k];
  return bad;
}

__global__ void __launch_bounds__(512, 1) k(long long *cyc, int *flag, int mode, int nw) {
  __shared__ __align__(16) double D[16][36];
  __shared__ __align__(16) double X[64][36];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < 16 * 36; i += 512) { const int e = i % 36; D[i / 36][e] = (e / 6 == e % 6) ? 40.0 + 0.01 * i : 0.3 + 1e-4 * i; }
  for (int i = tid; i < 64 * 36; i += 512) X[i / 36][i % 36] = 0.01 * (i % 97);
  __syncthreads();
  long long t0 = 0, t1 = 0;
  const int REP = 32;
  // (0) inverse6 in one lane, REP times back to back (dependent through shared memory)
  __syncthreads();
  t0 = clock64();
  if (mode == 0) {
    if (lane == 0 && warp < nw) for (int i = 0; i < REP; ++i) { if (block_inverse6(D[warp])) *flag = 1; }
  } else if (mode == 1) {
    if (lane == 0 && warp < nw) for (int i = 0; i < REP; ++i) { if (chol6(D[warp])) *flag = 1; D[warp][0] = 40.0; D[warp][7] = 41.0; D[warp][14] = 42; D[warp][21] = 43; D[warp][28] = 44; D[warp][35] = 45; }
  } else if (mode == 2) {
    for (int i = 0; i < REP; ++i) __syncthreads();
  } else if (mode == 3) {
    // panel: 6 lanes per row block: Y = X M
    const int g = lane / 6, r = lane - 6 * g;
    if (g < 5 && warp < nw) for (int i = 0; i < REP; ++i) {
      double2 *Dp = reinterpret_cast<double2 *>(&X[5 * warp % 64 + g][6 * r]);
      const double2 *M2 = reinterpret_cast<const double2 *>(D[warp]);
      const double2 v0 = Dp[0], v1 = Dp[1], v2 = Dp[2];
      double y[6];
#pragma unroll
      for (int c = 0; c < 6; ++c) { const double2 m0 = M2[3 * c], m1 = M2[3 * c + 1], m2 = M2[3 * c + 2]; y[c] = v0.x * m0.x + v0.y * m0.y + v1.x * m1.x + v1.y * m1.y + v2.x * m2.x + v2.y * m2.y; }
      Dp[0] = make_double2(y[0] * 1e-3, y[1] * 1e-3); Dp[1] = make_double2(y[2] * 1e-3, y[3] * 1e-3); Dp[2] = make_double2(y[4] * 1e-3, y[5] * 1e-3);
    }
  }
  t1 = clock64();
  __syncthreads();
  if (tid == 0) cyc[mode] = (clock64() - t0) / REP;
}
int main() {
  long long *cyc; int *flag;
  cudaMallocManaged(&cyc, 64 * 8); cudaMallocManaged(&flag, 4);
  for (int nw = 1; nw <= 16; nw *= 2) {
    for (int m = 0; m < 4; ++m) { k<<<1, 512>>>(cyc, flag, m, nw); cudaDeviceSynchronize(); }
    printf("%2d warps busy (one lane each): inverse6 %lld cycles, cholesky6 %lld cycles | __syncthreads(512) %lld | panel rows Y = X M (5 groups x 6 lanes) %lld | flag %d\n", nw, cyc[0], cyc[1], cyc[2], cyc[3], *flag);
  }
  // only warp 0 active
  return 0;
}
