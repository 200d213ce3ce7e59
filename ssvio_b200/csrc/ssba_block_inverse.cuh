// ssba_block_inverse.cuh — closed-form inverse of a symmetric positive definite 6x6 block (the pivot blocks of
// the block LDL^T factorisation in k_tree_solve, ssba_tree_program.hpp); host + device so that the CPU tests
// check the very same arithmetic.
#pragma once

#include <cuda_runtime.h>

#include "ssba_geometry.cuh"

namespace ssba {

SSBA_HD double block_rcp(double x) {
#ifdef __CUDA_ARCH__
  return __drcp_rn(x);
#else
  return 1.0 / x;
#endif
}

// adjugate and determinant of a symmetric 3x3 (a00 a10 a11 a20 a21 a22); returns false when a leading minor is <= 0 (or NaN)
SSBA_HD bool block_adj3(double a00, double a10, double a11, double a20, double a21, double a22, double *P, double &det) {
  P[0] = a11 * a22 - a21 * a21;   // 00
  P[1] = a20 * a21 - a10 * a22;   // 10
  P[2] = a00 * a22 - a20 * a20;   // 11
  P[3] = a10 * a21 - a20 * a11;   // 20
  P[4] = a10 * a20 - a00 * a21;   // 21
  P[5] = a00 * a11 - a10 * a10;   // 22
  det = a00 * P[0] + a10 * P[1] + a20 * P[3];
  return (a00 > 0.0) && (P[5] > 0.0) && (det > 0.0);
}

// M = D^-1 of the symmetric 6x6 block whose LOWER triangle is read from `D`, written back as the full symmetric
// 6x6, by one lane, in closed form: D = [[A, B^T], [B, C]] with 3x3 blocks, adjugates of A and of the Schur
// complement S = C - B A^-1 B^T (two reciprocals instead of a dependent chain of six square roots).  The
// leading minors a00, |A_2|, |A|, and s00, |S_2|, |S| are the running products of the Cholesky pivots, so
// "a pivot <= 0" (csparse_extension.cpp:115) is "a minor <= 0"; reported, and the block replaced by something
// finite (the trial is rejected anyway).
SSBA_HD bool block_inverse6(double *D) {
  const double2 r0 = *reinterpret_cast<const double2 *>(D);                                                  // a00
  const double2 r1 = *reinterpret_cast<const double2 *>(D + 6);                                              // a10 a11
  const double2 r2 = *reinterpret_cast<const double2 *>(D + 12), r2b = *reinterpret_cast<const double2 *>(D + 14);  // a20 a21 a22
  const double2 r3 = *reinterpret_cast<const double2 *>(D + 18), r3b = *reinterpret_cast<const double2 *>(D + 20);
  const double2 r4 = *reinterpret_cast<const double2 *>(D + 24), r4b = *reinterpret_cast<const double2 *>(D + 26), r4c = *reinterpret_cast<const double2 *>(D + 28);
  const double2 r5 = *reinterpret_cast<const double2 *>(D + 30), r5b = *reinterpret_cast<const double2 *>(D + 32), r5c = *reinterpret_cast<const double2 *>(D + 34);
  const double B[3][3] = {{r3.x, r3.y, r3b.x}, {r4.x, r4.y, r4b.x}, {r5.x, r5.y, r5b.x}};
  double P[6], det;
  bool ok = block_adj3(r0.x, r1.x, r1.y, r2.x, r2.y, r2b.x, P, det);
  const double Pf[3][3] = {{P[0], P[1], P[3]}, {P[1], P[2], P[4]}, {P[3], P[4], P[5]}};
  double T[3][3];  // B adj(A)
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int c = 0; c < 3; ++c) T[i][c] = B[i][0] * Pf[0][c] + B[i][1] * Pf[1][c] + B[i][2] * Pf[2][c];
  if (!ok) det = 1.0;
  const double rd = block_rcp(det);
  // S = C - (T B^T) / det, lower triangle
  const double u00 = T[0][0] * B[0][0] + T[0][1] * B[0][1] + T[0][2] * B[0][2];
  const double u10 = T[1][0] * B[0][0] + T[1][1] * B[0][1] + T[1][2] * B[0][2];
  const double u11 = T[1][0] * B[1][0] + T[1][1] * B[1][1] + T[1][2] * B[1][2];
  const double u20 = T[2][0] * B[0][0] + T[2][1] * B[0][1] + T[2][2] * B[0][2];
  const double u21 = T[2][0] * B[1][0] + T[2][1] * B[1][1] + T[2][2] * B[1][2];
  const double u22 = T[2][0] * B[2][0] + T[2][1] * B[2][1] + T[2][2] * B[2][2];
  const double s00 = r3b.y - u00 * rd, s10 = r4b.y - u10 * rd, s11 = r4c.x - u11 * rd;
  const double s20 = r5b.y - u20 * rd, s21 = r5c.x - u21 * rd, s22 = r5c.y - u22 * rd;
  double Q[6], dets;
  const bool ok2 = block_adj3(s00, s10, s11, s20, s21, s22, Q, dets);
  ok = ok && ok2;
  if (!ok2) dets = 1.0;
  const double rs = block_rcp(dets);
  const double N[3][3] = {{Q[0] * rs, Q[1] * rs, Q[3] * rs}, {Q[1] * rs, Q[2] * rs, Q[4] * rs}, {Q[3] * rs, Q[4] * rs, Q[5] * rs}};  // S^-1
  double G[3][3];  // B A^-1
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int c = 0; c < 3; ++c) G[i][c] = T[i][c] * rd;
  double M21[3][3];  // -S^-1 B A^-1
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int c = 0; c < 3; ++c) M21[i][c] = -(N[i][0] * G[0][c] + N[i][1] * G[1][c] + N[i][2] * G[2][c]);
  double M11[3][3];  // A^-1 - (B A^-1)^T M21
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int e = 0; e <= c; ++e) {
      const double v = Pf[c][e] * rd - (G[0][c] * M21[0][e] + G[1][c] * M21[1][e] + G[2][c] * M21[2][e]);
      M11[c][e] = v; M11[e][c] = v;
    }
  double2 *O = reinterpret_cast<double2 *>(D);
#pragma unroll
  for (int i = 0; i < 3; ++i) {  // rows 0..2: [M11 | M21^T]
    O[3 * i] = make_double2(M11[i][0], M11[i][1]);
    O[3 * i + 1] = make_double2(M11[i][2], M21[0][i]);
    O[3 * i + 2] = make_double2(M21[1][i], M21[2][i]);
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {  // rows 3..5: [M21 | S^-1]
    O[9 + 3 * i] = make_double2(M21[i][0], M21[i][1]);
    O[9 + 3 * i + 1] = make_double2(M21[i][2], N[i][0]);
    O[9 + 3 * i + 2] = make_double2(N[i][1], N[i][2]);
  }
  return !ok;
}

}  // namespace ssba
