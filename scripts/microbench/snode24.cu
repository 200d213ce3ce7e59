// Micro-benchmark of the supernode pieces of k_tree_solve (ssba_snode.cuh) on B200: in-register LDL^T of a 6w x 6w
// diagonal block by one warp, and the row solve of the rows below it; clock64 inside one CTA of 512 threads.
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
#include "snode.cuh"
using namespace ssba;

// block column layout of one supernode of width w in `pool`: column jb has (w - jb) diagonal-region blocks then m row blocks
__global__ void __launch_bounds__(512, 1) k(long long *cyc, double *out, int mode, int w, int nw) {
  extern __shared__ __align__(16) double pool[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = 6 * w, m = 8;
  int base[kSnMaxCols];
  int off = 0;
  for (int jb = 0; jb < kSnMaxCols; ++jb) { base[jb] = off; if (jb < w) off += 36 * (w - jb + m); }
  const int per_warp = off;
  double *mine = pool + per_warp * warp;
  // SPD: A = 30 I + 0.3 * pattern
  for (int e = lane; e < per_warp; e += 32) mine[e] = 0.05 + 1e-3 * (e % 13);
  __syncwarp();
  if (lane < 6) for (int jb = 0; jb < w; ++jb) mine[base[jb] + 7 * lane] = 30.0 + lane + jb;
  __syncthreads();
  const int REP = 16;
  long long t0 = clock64();
  double chk = 0.0;
  if (warp < nw) {
    if (mode == 0) {
      for (int rep = 0; rep < REP; ++rep) {
        double a[kSnMaxN];
        const int ib = lane / 6, ii = lane - 6 * ib;
#pragma unroll
        for (int jb = 0; jb < kSnMaxCols; ++jb) {
          if (jb < w && lane < n && jb <= ib) {
            const double2 *s = reinterpret_cast<const double2 *>(mine + base[jb] + 36 * (ib - jb) + 6 * ii);
            const double2 v0 = s[0], v1 = s[1], v2 = s[2];
            a[6 * jb] = v0.x; a[6 * jb + 1] = v0.y; a[6 * jb + 2] = v1.x; a[6 * jb + 3] = v1.y; a[6 * jb + 4] = v2.x; a[6 * jb + 5] = v2.y;
          } else {
#pragma unroll
            for (int jj = 0; jj < 6; ++jj) a[6 * jb + jj] = 0.0;
          }
        }
        const bool bad = sn_factor(a, n, lane);
        if (bad) chk += 1e9;
        chk += a[0] + a[7] + a[23];
        // restore-ish: write something back so that the repetitions depend on each other through shared memory
        if (lane < n) mine[base[0] + 6 * lane % 36] += 1e-12 * a[0];
        __syncwarp();
      }
    } else {
      // rows below: 30 lanes = 5 row blocks of the m
      for (int rep = 0; rep < REP; ++rep) {
        const int g = lane / 6, r = lane - 6 * g;
        double x[kSnMaxN], y[kSnMaxN];
#pragma unroll
        for (int jb = 0; jb < kSnMaxCols; ++jb) {
          if (jb < w && g < 5) {
            const double2 *s = reinterpret_cast<const double2 *>(mine + base[jb] + 36 * (w - jb + g) + 6 * r);
            const double2 v0 = s[0], v1 = s[1], v2 = s[2];
            x[6 * jb] = v0.x; x[6 * jb + 1] = v0.y; x[6 * jb + 2] = v1.x; x[6 * jb + 3] = v1.y; x[6 * jb + 4] = v2.x; x[6 * jb + 5] = v2.y;
          } else {
#pragma unroll
            for (int jj = 0; jj < 6; ++jj) x[6 * jb + jj] = 0.0;
          }
        }
        if (g < 5) {
          sn_row_solve(x, y, n, mine, base);
#pragma unroll
          for (int jb = 0; jb < kSnMaxCols; ++jb) {
            if (jb < w) {
              double2 *s = reinterpret_cast<double2 *>(mine + base[jb] + 36 * (w - jb + g) + 6 * r);
              s[0] = make_double2(1e-3 * y[6 * jb], 1e-3 * y[6 * jb + 1]); s[1] = make_double2(1e-3 * y[6 * jb + 2], 1e-3 * y[6 * jb + 3]); s[2] = make_double2(1e-3 * y[6 * jb + 4], 1e-3 * y[6 * jb + 5]);
            }
          }
        }
        chk += x[3];
        __syncwarp();
      }
    }
  }
  const long long t1 = clock64();
  if (lane == 0 && warp == 0) cyc[0] = (t1 - t0) / REP;
  out[tid] = chk;
}

int main() {
  long long *cyc; double *out;
  cudaMallocManaged(&cyc, 64); cudaMallocManaged(&out, 512 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int w = 1; w <= 4; ++w)
    for (int nw = 1; nw <= 16; nw *= 4) {
      long long c[2];
      for (int mode = 0; mode < 2; ++mode) { k<<<1, 512, 200 * 1024>>>(cyc, out, mode, w, nw); cudaDeviceSynchronize(); c[mode] = cyc[0]; }
      printf("width %d (%2d x %2d), %2d warps busy: factor %lld cycles | row solve of 30 rows %lld cycles | %s\n", w, 6 * w, 6 * w, nw, c[0], c[1], cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
