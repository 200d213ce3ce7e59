// ssba_snode.cuh — dense pieces of a supernode (up to four 6x6 block columns with the same row pattern) of the
// reduced pose solve, one scalar row per lane, everything in registers (k_tree_solve, ssba_tree_solve.cu).
// Numerics: the scalar LDL^T of the supernode's diagonal block - the pivots d_k are the squares of CSparse's
// Cholesky pivots (cs_chol, g2o/solvers/csparse/csparse_extension.cpp:95-118), "d_k <= 0 => not positive definite"
// is the same test (:115).
#pragma once

#include "../../ssvio_b200/csrc/ssba_block_inverse.cuh"

namespace ssba {

constexpr int kSnMaxCols = 4;              // block columns of a supernode
constexpr int kSnMaxN = 6 * kSnMaxCols;    // scalar columns

// Lane i < n holds row i of the symmetric n x n block: a[j] = A[i][j], j <= i (the rest is ignored).
// Returns with a[j] = L[i][j] (j < i) and a[i] = 1 / d_i on lane i; the result is true when a pivot was not positive.
// All 32 lanes must call (shuffles); n is a multiple of 6, warp-uniform.
__device__ __forceinline__ bool sn_factor(double (&a)[kSnMaxN], int n, int lane) {
  bool bad = false;
#pragma unroll
  for (int kb = 0; kb < kSnMaxCols; ++kb) {
    if (6 * kb < n) {
#pragma unroll
      for (int kk = 0; kk < 6; ++kk) {
        const int k = 6 * kb + kk;
        const double ck = a[k];  // A[i][k] as updated by the columns before k
        const double dk = __shfl_sync(0xffffffffu, ck, k);
        if (!(dk > 0.0)) bad = true;
#if defined(SN_RCP_NONE)
        const double inv = dk * 1e-3;
#elif defined(SN_RCP_FAST)
        double inv;
        {
          double x0;
          asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x0) : "d"(dk));
          const double e0 = fma(-dk, x0, 1.0);
          const double x1 = fma(x0, e0, x0);
          const double e1 = fma(-dk, x1, 1.0);
          inv = fma(x1, e1, x1);
        }
#else
        const double inv = block_rcp(dk);
#endif
        const double lik = ck * inv;
#pragma unroll
        for (int j = k + 1; j < 6 * (kb + 1); ++j) a[j] -= lik * __shfl_sync(0xffffffffu, ck, j);
#pragma unroll
        for (int jb = kb + 1; jb < kSnMaxCols; ++jb) {
          if (6 * jb < n) {
#pragma unroll
            for (int jj = 0; jj < 6; ++jj) a[6 * jb + jj] -= lik * __shfl_sync(0xffffffffu, ck, 6 * jb + jj);
          }
        }
        a[k] = lane == k ? inv : lik;
      }
    }
  }
  return bad;
}

// One scalar row of the rows below a factored supernode: x L^T = a by forward substitution (x = the unscaled row
// X = L21 D of the block LDL^T), then y = x D^-1.  `Lrow(k)` = shared-memory address of row k of the factor as the
// block column layout stores it: Lrow(k)[j - 6 * (j / 6) + 36-double stride per block column] - the caller passes the
// four block-column bases of the supernode; row k of block row kb, column block jb sits at
// base[jb] + 36 * (kb - jb) + 6 * kk.  On return x holds the unscaled row and y the scaled one.
__device__ __forceinline__ void sn_row_solve(double (&x)[kSnMaxN], double (&y)[kSnMaxN], int n, const double *pool, const int (&base)[kSnMaxCols]) {
#pragma unroll
  for (int kb = 0; kb < kSnMaxCols; ++kb) {
    if (6 * kb < n) {
#pragma unroll
      for (int kk = 0; kk < 6; ++kk) {
        const int k = 6 * kb + kk;
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int jb = 0; jb < kb; ++jb) {
          const double2 *L2 = reinterpret_cast<const double2 *>(pool + base[jb] + 36 * (kb - jb) + 6 * kk);
          const double2 l0 = L2[0], l1 = L2[1], l2 = L2[2];
          s0 += x[6 * jb] * l0.x; s1 += x[6 * jb + 1] * l0.y;
          s0 += x[6 * jb + 2] * l1.x; s1 += x[6 * jb + 3] * l1.y;
          s0 += x[6 * jb + 4] * l2.x; s1 += x[6 * jb + 5] * l2.y;
        }
        const double *Ld = pool + base[kb] + 6 * kk;  // the diagonal block's row: L[k][6 kb .. k), 1 / d_k at [kk]
#pragma unroll
        for (int jj = 0; jj < kk; ++jj) { if (jj & 1) s1 += x[6 * kb + jj] * Ld[jj]; else s0 += x[6 * kb + jj] * Ld[jj]; }
        x[k] -= s0 + s1;
        y[k] = x[k] * Ld[kk];
      }
    }
  }
}

}  // namespace ssba
