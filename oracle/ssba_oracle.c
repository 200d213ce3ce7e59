/*
 * ssba_oracle.c — plain-C restatement of ssvio's back-end local bundle adjustment hot path.
 *
 * TEST INFRASTRUCTURE (parity checker / CPU baseline "port").  Never linked into, loaded by or
 * used as a fallback for the product library libssba.so.  See ssba_oracle.h for how it is pinned
 * against the compiled reference (oracle/_ref).
 *
 * What is restated (paths relative to the ssvio tree; g2o/ = thirdparty/g2o/g2o/):
 *   graph semantics      src/ssvio/backend.cpp:81-178
 *   vertex / edge types  include/ssvio/g2otypes.hpp:28-65,112-131
 *   SE(3) arithmetic     thirdparty/sophus/sophus/so3.hpp, se3.hpp (lines cited per function)
 *   active sets, order   g2o/core/sparse_optimizer.cpp:168-272
 *   numeric Jacobians    g2o/core/base_binary_edge.hpp:144-212
 *   quadratic form       g2o/core/base_binary_edge.hpp:61-134, robust_kernel_impl.cpp:65-78
 *   build / damp / Schur g2o/core/block_solver.hpp:314-565
 *   LM control           g2o/core/optimization_algorithm_levenberg.cpp:58-175
 *   outer loop           g2o/core/sparse_optimizer.cpp:366-431
 * The reduced pose system is solved by a natural-order envelope (skyline) Cholesky instead of
 * CSparse's AMD-ordered up-looking Cholesky (g2o/solvers/csparse/csparse_extension.cpp:67-122):
 * the same factorisation up to the elimination order, i.e. up to rounding; the non-positive
 * pivot => failure rule (:115) is kept.
 */
#define _POSIX_C_SOURCE 200809L
#include "ssba_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ------------------------------------------------------------------ small fixed-size algebra */

static void cross3(const double a[3], const double b[3], double o[3]) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}

/* Sophus SO3 point action (so3.hpp:352-360): p + w*2(v x p) + v x (2(v x p)); q = (x,y,z,w) */
static void quat_rotate(const double q[4], const double p[3], double o[3]) {
  double uv[3], c[3];
  cross3(q, p, uv);
  uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
  cross3(q, uv, c);
  o[0] = p[0] + q[3] * uv[0] + c[0];
  o[1] = p[1] + q[3] * uv[1] + c[1];
  o[2] = p[2] + q[3] * uv[2] + c[2];
}

/* Sophus SO3 product (so3.hpp:322-334) followed by the re-normalisation done by the
 * SO3(quaternion) constructor (so3.hpp:498-503 -> normalize() :293-298). */
static void quat_mul_normalized(const double a[4], const double b[4], double o[4]) {
  const double ax = a[0], ay = a[1], az = a[2], aw = a[3];
  const double bx = b[0], by = b[1], bz = b[2], bw = b[3];
  double w = aw * bw - ax * bx - ay * by - az * bz;
  double x = aw * bx + ax * bw + ay * bz - az * by;
  double y = aw * by + ay * bw + az * bx - ax * bz;
  double z = aw * bz + az * bw + ax * by - ay * bx;
  double len = sqrt(x * x + y * y + z * z + w * w);
  o[0] = x / len; o[1] = y / len; o[2] = z / len; o[3] = w / len;
}

/* rotation matrix of a unit quaternion (Eigen::QuaternionBase::toRotationMatrix), row-major */
static void quat_to_matrix(const double q[4], double R[9]) {
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w;
  const double txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
  R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}

/* SE3::exp (se3.hpp:763-784) with SO3::expAndTheta (so3.hpp:593-622); a = (upsilon, omega) */
void ssba_oracle_se3_exp(const double a[6], double qt[7]) {
  const double *ups = a, *om = a + 3;
  const double eps = 1e-10; /* Sophus::Constants<double>::epsilon(), common.hpp:111 */
  double theta_sq = om[0] * om[0] + om[1] * om[1] + om[2] * om[2];
  double theta, imag, real;
  if (theta_sq < eps * eps) {
    double theta_po4 = theta_sq * theta_sq;
    theta = 0.0;
    imag = 0.5 - (1.0 / 48.0) * theta_sq + (1.0 / 3840.0) * theta_po4;
    real = 1.0 - (1.0 / 8.0) * theta_sq + (1.0 / 384.0) * theta_po4;
  } else {
    double half;
    theta = sqrt(theta_sq);
    half = 0.5 * theta;
    imag = sin(half) / theta;
    real = cos(half);
  }
  qt[0] = imag * om[0]; qt[1] = imag * om[1]; qt[2] = imag * om[2]; qt[3] = real;

  /* V * upsilon */
  double V[9];
  if (theta < eps) {
    quat_to_matrix(qt, V); /* V = so3.matrix() (se3.hpp:774-776) */
  } else {
    /* V = I + (1-cos)/theta^2 * Omega + (theta - sin)/(theta^2*theta) * Omega^2 (:777-782) */
    double O[9] = {0, -om[2], om[1], om[2], 0, -om[0], -om[1], om[0], 0};
    double O2[9];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c)
        O2[3 * r + c] = O[3 * r] * O[c] + O[3 * r + 1] * O[3 + c] + O[3 * r + 2] * O[6 + c];
    double th2 = theta * theta;
    double c1 = (1.0 - cos(theta)) / th2;
    double c2 = (theta - sin(theta)) / (th2 * theta);
    for (int i = 0; i < 9; ++i) V[i] = c1 * O[i] + c2 * O2[i];
    V[0] += 1.0; V[4] += 1.0; V[8] += 1.0;
  }
  for (int r = 0; r < 3; ++r)
    qt[4 + r] = V[3 * r] * ups[0] + V[3 * r + 1] * ups[1] + V[3 * r + 2] * ups[2];
}

/* SE3 product (se3.hpp:309-314): (qa*qb normalised, ta + qa*tb) */
void ssba_oracle_se3_mul(const double a[7], const double b[7], double out[7]) {
  double q[4], rt[3];
  quat_mul_normalized(a, b, q);
  quat_rotate(a, b + 4, rt);
  out[0] = q[0]; out[1] = q[1]; out[2] = q[2]; out[3] = q[3];
  out[4] = a[4] + rt[0]; out[5] = a[5] + rt[1]; out[6] = a[6] + rt[2];
}

/* VertexPose::oplusImpl (g2otypes.hpp:36-41): T <- exp(delta) * T */
void ssba_oracle_pose_oplus(const double qt[7], const double delta[6], double out[7]) {
  double e[7];
  ssba_oracle_se3_exp(delta, e);
  ssba_oracle_se3_mul(e, qt, out);
}

/* SE3 point action (se3.hpp:325-328) */
static void se3_act(const double qt[7], const double p[3], double o[3]) {
  quat_rotate(qt, p, o);
  o[0] += qt[4]; o[1] += qt[5]; o[2] += qt[6];
}

/* EdgeProjection::computeError (g2otypes.hpp:123-131):
 * e = z - (K * (ext * (T * p))) / depth, no cheirality test, no epsilon on the depth. */
void ssba_oracle_edge_error(const double K[9], const double ext[7], const double pose[7],
                            const double p[3], const double uv[2], double err[2]) {
  double pb[3], pc[3], px[3];
  se3_act(pose, p, pb);
  se3_act(ext, pb, pc);
  for (int r = 0; r < 3; ++r) px[r] = K[3 * r] * pc[0] + K[3 * r + 1] * pc[1] + K[3 * r + 2] * pc[2];
  err[0] = uv[0] - px[0] / px[2];
  err[1] = uv[1] - px[1] / px[2];
}

/* Jacobians of the error w.r.t. the pose tangent (2x6, row-major) and the point (2x3).
 * mode 1: central differences, delta = 1e-9 (base_binary_edge.hpp:144-212) — as shipped,
 *         because the analytic override is commented out (g2otypes.hpp:133-153).
 * mode 0: closed form  J_xi = Jpi R_ext [I, -[T p]x],  J_p = Jpi R_ext R  (SURVEY.md 8a a4). */
void ssba_oracle_edge_jacobians(const double K[9], const double ext[7], const double pose[7],
                                const double p[3], const double uv[2], int32_t mode,
                                double Jx[12], double Jp[6]) {
  if (mode == SSBA_JACOBIAN_NUMERIC) {
    const double delta = 1e-9;
    const double scalar = 1 / (2 * delta);
    for (int d = 0; d < 6; ++d) {
      double add[6] = {0, 0, 0, 0, 0, 0}, T1[7], ep[2], em[2];
      add[d] = delta;
      ssba_oracle_pose_oplus(pose, add, T1);
      ssba_oracle_edge_error(K, ext, T1, p, uv, ep);
      add[d] = -delta;
      ssba_oracle_pose_oplus(pose, add, T1);
      ssba_oracle_edge_error(K, ext, T1, p, uv, em);
      Jx[d] = scalar * (ep[0] - em[0]);
      Jx[6 + d] = scalar * (ep[1] - em[1]);
    }
    for (int d = 0; d < 3; ++d) {
      double p1[3] = {p[0], p[1], p[2]}, ep[2], em[2];
      p1[d] = p[d] + delta;
      ssba_oracle_edge_error(K, ext, pose, p1, uv, ep);
      p1[d] = p[d] + (-delta);
      ssba_oracle_edge_error(K, ext, pose, p1, uv, em);
      Jp[d] = scalar * (ep[0] - em[0]);
      Jp[3 + d] = scalar * (ep[1] - em[1]);
    }
    return;
  }
  double pb[3], pc[3], R[9], Re[9], Jpi[6], JR[6];
  se3_act(pose, p, pb);
  se3_act(ext, pb, pc);
  quat_to_matrix(pose, R);
  quat_to_matrix(ext, Re);
  /* general K: u = (K0.pc)/(K2.pc); for the pin-hole K of the reference this is
   * -[[fx/Z, 0, -fx X/Z^2], [0, fy/Z, -fy Y/Z^2]] */
  double n0 = K[0] * pc[0] + K[1] * pc[1] + K[2] * pc[2];
  double n1 = K[3] * pc[0] + K[4] * pc[1] + K[5] * pc[2];
  double dn = K[6] * pc[0] + K[7] * pc[1] + K[8] * pc[2];
  double id = 1.0 / dn, id2 = id * id;
  for (int c = 0; c < 3; ++c) {
    Jpi[c] = -(K[c] * id - n0 * K[6 + c] * id2);
    Jpi[3 + c] = -(K[3 + c] * id - n1 * K[6 + c] * id2);
  }
  for (int r = 0; r < 2; ++r)
    for (int c = 0; c < 3; ++c)
      JR[3 * r + c] = Jpi[3 * r] * Re[c] + Jpi[3 * r + 1] * Re[3 + c] + Jpi[3 * r + 2] * Re[6 + c];
  for (int r = 0; r < 2; ++r) {
    const double *j = JR + 3 * r;
    Jx[6 * r + 0] = j[0]; Jx[6 * r + 1] = j[1]; Jx[6 * r + 2] = j[2];
    /* -JR * hat(pb) */
    Jx[6 * r + 3] = -(j[1] * pb[2] - j[2] * pb[1]);
    Jx[6 * r + 4] = -(j[2] * pb[0] - j[0] * pb[2]);
    Jx[6 * r + 5] = -(j[0] * pb[1] - j[1] * pb[0]);
    for (int c = 0; c < 3; ++c)
      Jp[3 * r + c] = j[0] * R[c] + j[1] * R[3 + c] + j[2] * R[6 + c];
  }
}

/* RobustKernelHuber::robustify (robust_kernel_impl.cpp:65-78); delta <= 0: no kernel */
void ssba_oracle_huber(double e, double delta, double rho[3]) {
  double dsqr = delta * delta;
  if (delta <= 0 || e <= dsqr) {
    rho[0] = e; rho[1] = 1.; rho[2] = 0.;
  } else {
    double sqrte = sqrt(e);
    rho[0] = 2 * sqrte * delta - dsqr;
    rho[1] = delta / sqrte;
    rho[2] = -0.5 * rho[1] / e;
  }
}

/* Eigen 3.3.7 closed-form 3x3 inverse (cofactors / determinant) used at block_solver.hpp:350.
 * m is symmetric 3x3 row-major. Returns inverse in o. */
static void inv3(const double m[9], double o[9]) {
  double c00 = m[4] * m[8] - m[5] * m[7];
  double c10 = m[5] * m[6] - m[3] * m[8];
  double c20 = m[3] * m[7] - m[4] * m[6];
  double det = m[0] * c00 + m[1] * c10 + m[2] * c20;
  double id = 1.0 / det;
  o[0] = c00 * id;
  o[3] = c10 * id;
  o[6] = c20 * id;
  o[1] = (m[2] * m[7] - m[1] * m[8]) * id;
  o[4] = (m[0] * m[8] - m[2] * m[6]) * id;
  o[7] = (m[1] * m[6] - m[0] * m[7]) * id;
  o[2] = (m[1] * m[5] - m[2] * m[4]) * id;
  o[5] = (m[2] * m[3] - m[0] * m[5]) * id;
  o[8] = (m[0] * m[4] - m[1] * m[3]) * id;
}

/* ------------------------------------------------------------------ problem state */

typedef struct {
  /* inputs */
  const double *K, *ext, *uv;
  const int32_t *epose, *epoint;
  const uint8_t *ecam;
  int n_poses, n_points, n_edges;
  double huber;
  int jac_mode;
  /* estimates */
  double *pose, *point;   /* current */
  double *pose_bak, *point_bak;
  /* active sets (sparse_optimizer.cpp:201-272, 168-192) */
  int *pose_h, *point_h;  /* hessian index among free poses / free landmarks, -1 = fixed/inactive */
  int *active_edge;       /* list of active edge ids in addEdge order */
  int n_active;
  int n_fp, n_fl;         /* free poses, free landmarks */
  int *fp_vertex, *fl_vertex;
  /* (pose, landmark) pairs = Hpl blocks (block_solver.hpp:181-211) */
  int *edge_pair;         /* per active edge: pair id or -1 */
  int n_pairs;
  int *pair_pose, *pair_lm;   /* free indices */
  int *lm_pair_ptr;       /* CSR over free landmarks -> pairs sorted by pose */
  /* system */
  double *Hpp;  /* n_fp x 36 row-major */
  double *Hll;  /* n_fl x 9 */
  double *Hpl;  /* n_pairs x 18 (6x3 row-major) */
  double *b;    /* 6 n_fp + 3 n_fl */
  double *x;
  double *err;  /* n_edges x 2 */
  double *Dinv; /* n_fl x 9 */
  double *S;    /* (6 n_fp)^2 dense, row-major, lower triangle used by the factorisation */
  int *first;   /* envelope: first non-zero column per row */
  double *bschur, *coeff;
  /* landmark shard filter (multi-GPU parity tests): point row -> owning rank, NULL = all */
  const int32_t *owner;
  int rank;
  uint8_t *cam0;
} Prob;

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* SparseOptimizer::computeActiveErrors (sparse_optimizer.cpp:63-90) */
static void compute_active_errors(Prob *P) {
  for (int k = 0; k < P->n_active; ++k) {
    int e = P->active_edge[k];
    ssba_oracle_edge_error(P->K, P->ext + 7 * P->ecam[e], P->pose + 7 * P->epose[e],
                           P->point + 3 * P->epoint[e], P->uv + 2 * e, P->err + 2 * e);
  }
}

/* activeRobustChi2 / activeChi2 (sparse_optimizer.cpp:92-116): serial sums in edge order.
 * chi2 = e.dot(I * e) (base_edge.h:79-82, information = identity backend.cpp:161) */
static double active_chi2(const Prob *P, int robust) {
  double chi = 0.0, rho[3];
  for (int k = 0; k < P->n_active; ++k) {
    const double *e = P->err + 2 * P->active_edge[k];
    double c = e[0] * e[0] + e[1] * e[1];
    if (robust) { ssba_oracle_huber(c, P->huber, rho); chi += rho[0]; }
    else chi += c;
  }
  return chi;
}

/* BlockSolver::buildSystem (block_solver.hpp:462-521): linearizeOplus + constructQuadraticForm
 * per active edge in order (base_binary_edge.hpp:61-134). */
static void build_system(Prob *P) {
  memset(P->Hpp, 0, sizeof(double) * 36 * P->n_fp);
  memset(P->Hll, 0, sizeof(double) * 9 * P->n_fl);
  memset(P->Hpl, 0, sizeof(double) * 18 * P->n_pairs);
  memset(P->b, 0, sizeof(double) * (6 * P->n_fp + 3 * P->n_fl));
  for (int k = 0; k < P->n_active; ++k) {
    int e = P->active_edge[k];
    if (P->owner && P->owner[P->epoint[e]] != P->rank) continue;
    int ip = P->pose_h[P->epose[e]], il = P->point_h[P->epoint[e]];
    double Jx[12], Jp[6], rho[3];
    ssba_oracle_edge_jacobians(P->K, P->ext + 7 * P->ecam[e], P->pose + 7 * P->epose[e],
                               P->point + 3 * P->epoint[e], P->uv + 2 * e, P->jac_mode, Jx, Jp);
    const double *er = P->err + 2 * e;
    double chi = er[0] * er[0] + er[1] * er[1];
    ssba_oracle_huber(chi, P->huber, rho);
    double w = rho[1];
    double r0 = -er[0] * w, r1 = -er[1] * w; /* omega_r = -Omega e, times rho' */
    if (ip >= 0) {
      double *bp = P->b + 6 * ip, *H = P->Hpp + 36 * ip;
      for (int a = 0; a < 6; ++a) {
        bp[a] += Jx[a] * r0 + Jx[6 + a] * r1;
        for (int c = 0; c < 6; ++c) H[6 * a + c] += w * (Jx[a] * Jx[c] + Jx[6 + a] * Jx[6 + c]);
      }
      if (il >= 0) {
        double *W = P->Hpl + 18 * P->edge_pair[k];
        for (int a = 0; a < 6; ++a)
          for (int c = 0; c < 3; ++c) W[3 * a + c] += w * (Jx[a] * Jp[c] + Jx[6 + a] * Jp[3 + c]);
      }
    }
    if (il >= 0) {
      double *bl = P->b + 6 * P->n_fp + 3 * il, *H = P->Hll + 9 * il;
      for (int a = 0; a < 3; ++a) {
        bl[a] += Jp[a] * r0 + Jp[3 + a] * r1;
        for (int c = 0; c < 3; ++c) H[3 * a + c] += w * (Jp[a] * Jp[c] + Jp[3 + a] * Jp[3 + c]);
      }
    }
  }
}

/* computeLambdaInit (levenberg.cpp:152-166) */
static double lambda_init(const Prob *P, double tau) {
  double m = 0;
  for (int i = 0; i < P->n_fp; ++i)
    for (int d = 0; d < 6; ++d) m = fmax(fabs(P->Hpp[36 * i + 7 * d]), m);
  for (int j = 0; j < P->n_fl; ++j)
    for (int d = 0; d < 3; ++d) m = fmax(fabs(P->Hll[9 * j + 4 * d]), m);
  return tau * m;
}

/* envelope Cholesky of the lower triangle of S (n x n, row-major), in place: S = L L^T.
 * Returns 0 when a pivot is <= 0 (csparse_extension.cpp:115), 1 otherwise. */
static int skyline_cholesky(double *S, int n, const int *first) {
  for (int i = 0; i < n; ++i) {
    double *Li = S + (size_t)i * n;
    for (int j = first[i]; j <= i; ++j) {
      const double *Lj = S + (size_t)j * n;
      int k0 = first[i] > first[j] ? first[i] : first[j];
      double s = Li[j];
      for (int k = k0; k < j; ++k) s -= Li[k] * Lj[k];
      if (j < i) {
        Li[j] = s / Lj[j];
      } else {
        if (s <= 0) return 0;
        Li[i] = sqrt(s);
      }
    }
  }
  return 1;
}

static void skyline_solve(const double *L, int n, const int *first, double *x) {
  for (int i = 0; i < n; ++i) { /* L y = b */
    const double *Li = L + (size_t)i * n;
    double s = x[i];
    for (int k = first[i]; k < i; ++k) s -= Li[k] * x[k];
    x[i] = s / Li[i];
  }
  for (int i = n - 1; i >= 0; --i) { /* L^T x = y */
    const double *Li = L + (size_t)i * n;
    x[i] /= Li[i];
    for (int k = first[i]; k < i; ++k) x[k] -= Li[k] * x[i];
  }
}

/* The reduced (Schur) system of BlockSolver::solve (block_solver.hpp:334-400): upper block
 * triangle of S = (Hpp + lambda I) - sum_l W Dinv W^T in P->S, bschur in P->bschur.  With a
 * shard filter only the owned landmarks contribute and only rank 0 adds lambda, so that the
 * per-rank systems add up to the full one. */
static void assemble_reduced_system(Prob *P, double lambda) {
  const int np = P->n_fp, nl = P->n_fl, n = 6 * np;
  double *bl = P->b + n;
  const double lam_p = (!P->owner || P->rank == 0) ? lambda : 0.0;
  /* Hschur = Hpp + lambda I, pattern of the Schur complement (:334-335, :534-539) */
  memset(P->S, 0, sizeof(double) * (size_t)n * n);
  for (int i = 0; i < np; ++i)
    for (int r = 0; r < 6; ++r)
      for (int c = 0; c < 6; ++c)
        P->S[(size_t)(6 * i + r) * n + 6 * i + c] = P->Hpp[36 * i + 6 * r + c] + (r == c ? lam_p : 0.0);
  memset(P->coeff, 0, sizeof(double) * n);
  for (int l = 0; l < nl; ++l) { /* :342-393 */
    if (P->owner && P->owner[P->fl_vertex[l]] != P->rank) continue;
    double D[9], *Di = P->Dinv + 9 * l, db[3];
    memcpy(D, P->Hll + 9 * l, sizeof(D));
    D[0] += lambda; D[4] += lambda; D[8] += lambda; /* :541-548 */
    inv3(D, Di);
    for (int r = 0; r < 3; ++r) db[r] = Di[3 * r] * bl[3 * l] + Di[3 * r + 1] * bl[3 * l + 1] + Di[3 * r + 2] * bl[3 * l + 2];
    for (int a = P->lm_pair_ptr[l]; a < P->lm_pair_ptr[l + 1]; ++a) {
      const double *Wi = P->Hpl + 18 * a;
      int i1 = P->pair_pose[a];
      double BD[18];
      for (int r = 0; r < 6; ++r)
        for (int c = 0; c < 3; ++c)
          BD[3 * r + c] = Wi[3 * r] * Di[c] + Wi[3 * r + 1] * Di[3 + c] + Wi[3 * r + 2] * Di[6 + c];
      for (int r = 0; r < 6; ++r)
        P->coeff[6 * i1 + r] += Wi[3 * r] * db[0] + Wi[3 * r + 1] * db[1] + Wi[3 * r + 2] * db[2];
      for (int a2 = a; a2 < P->lm_pair_ptr[l + 1]; ++a2) { /* i2 >= i1, upper blocks (:380-391) */
        const double *Wj = P->Hpl + 18 * a2;
        int i2 = P->pair_pose[a2];
        for (int r = 0; r < 6; ++r)
          for (int c = 0; c < 6; ++c)
            P->S[(size_t)(6 * i1 + r) * n + 6 * i2 + c] -=
                BD[3 * r] * Wj[3 * c] + BD[3 * r + 1] * Wj[3 * c + 1] + BD[3 * r + 2] * Wj[3 * c + 2];
      }
    }
  }
  for (int i = 0; i < n; ++i) P->bschur[i] = P->b[i] - P->coeff[i]; /* :397-400 */
}

/* BlockSolver::solve with setLambda/restoreDiagonal folded in (block_solver.hpp:314-447,524-565).
 * Returns 0 when the reduced Cholesky fails. */
static int solve_damped(Prob *P, double lambda) {
  const int np = P->n_fp, nl = P->n_fl, n = 6 * np;
  double *bl = P->b + n;
  (void)np;
  assemble_reduced_system(P, lambda);
  /* mirror the upper block triangle into the lower one and take the envelope */
  for (int i = 0; i < n; ++i) {
    int f = i;
    for (int j = 0; j < i; ++j) {
      double v = P->S[(size_t)j * n + i];
      P->S[(size_t)i * n + j] = v;
      if (v != 0.0 && j < f) f = j;
    }
    /* envelope at block granularity so structural zeros inside a block stay in */
    P->first[i] = (f / 6) * 6;
  }
  /* LinearSolverCSparse::solve copies b into x before factorising
   * (linear_solver_csparse.h:122-124), so a failed factorisation leaves x_p = bschur and the
   * landmark part of x stale; the trial is then rejected through tempChi = DBL_MAX. */
  memcpy(P->x, P->bschur, sizeof(double) * n);
  if (!skyline_cholesky(P->S, n, P->first)) return 0;
  skyline_solve(P->S, n, P->first, P->x);
  /* landmarks: x_l = Dinv (b_l - W^T x_p) (:420-442) */
  for (int l = 0; l < nl; ++l) {
    double c[3] = {bl[3 * l], bl[3 * l + 1], bl[3 * l + 2]};
    for (int a = P->lm_pair_ptr[l]; a < P->lm_pair_ptr[l + 1]; ++a) {
      const double *W = P->Hpl + 18 * a;
      const double *xp = P->x + 6 * P->pair_pose[a];
      for (int r = 0; r < 6; ++r) {
        c[0] -= W[3 * r] * xp[r]; c[1] -= W[3 * r + 1] * xp[r]; c[2] -= W[3 * r + 2] * xp[r];
      }
    }
    const double *Di = P->Dinv + 9 * l;
    for (int r = 0; r < 3; ++r) P->x[n + 3 * l + r] = Di[3 * r] * c[0] + Di[3 * r + 1] * c[1] + Di[3 * r + 2] * c[2];
  }
  return 1;
}

/* SparseOptimizer::update (sparse_optimizer.cpp:433-446) */
static void apply_update(Prob *P) {
  for (int i = 0; i < P->n_fp; ++i) {
    double out[7];
    double *T = P->pose + 7 * P->fp_vertex[i];
    ssba_oracle_pose_oplus(T, P->x + 6 * i, out);
    memcpy(T, out, sizeof(out));
  }
  const double *xl = P->x + 6 * P->n_fp;
  for (int j = 0; j < P->n_fl; ++j) {
    double *p = P->point + 3 * P->fl_vertex[j];
    p[0] += xl[3 * j]; p[1] += xl[3 * j + 1]; p[2] += xl[3 * j + 2];
  }
}

static int cmp_ll(const void *a, const void *b) {
  long long x = *(const long long *)a, y = *(const long long *)b;
  return (x > y) - (x < y);
}

/* Graph -> problem: active sets, index mapping, Hpl pattern, work arrays.  Returns 1 when there
 * is nothing to optimise (optimize() would return -1), 0 otherwise. */
static int prob_setup(Prob *P, const double K[9], const double *ext_qt, int32_t n_poses,
                      const double *poses_qt, const uint8_t *pose_fixed, int32_t n_points,
                      const double *points, const uint8_t *point_fixed, int32_t n_edges,
                      const int32_t *pose_idx, const int32_t *point_idx, const uint8_t *cam_idx,
                      const double *uv, double huber_delta, int32_t jacobian_mode) {
  memset(P, 0, sizeof(*P));
  P->K = K; P->ext = ext_qt; P->uv = uv; P->epose = pose_idx; P->epoint = point_idx;
  P->n_poses = n_poses; P->n_points = n_points; P->n_edges = n_edges;
  P->huber = huber_delta; P->jac_mode = jacobian_mode;
  if (!cam_idx) { P->cam0 = calloc(n_edges > 0 ? n_edges : 1, 1); P->ecam = P->cam0; } else P->ecam = cam_idx;

  P->pose = malloc(sizeof(double) * 7 * (n_poses + 1));
  P->point = malloc(sizeof(double) * 3 * (n_points + 1));
  P->pose_bak = malloc(sizeof(double) * 7 * (n_poses + 1));
  P->point_bak = malloc(sizeof(double) * 3 * (n_points + 1));
  memcpy(P->pose, poses_qt, sizeof(double) * 7 * n_poses);
  memcpy(P->point, points, sizeof(double) * 3 * n_points);

  /* initializeOptimization (sparse_optimizer.cpp:201-272): an edge is active unless both its
   * vertices are fixed (:237); a vertex is active when it has an active edge; index mapping
   * (:168-192): free poses in id order, then free (marginalised) landmarks in id order. */
  P->active_edge = malloc(sizeof(int) * (n_edges + 1));
  char *pa = calloc(n_poses + 1, 1), *la = calloc(n_points + 1, 1);
  for (int e = 0; e < n_edges; ++e) {
    int pf = pose_fixed && pose_fixed[pose_idx[e]], lf = point_fixed && point_fixed[point_idx[e]];
    if (pf && lf) continue;
    P->active_edge[P->n_active++] = e;
    pa[pose_idx[e]] = 1; la[point_idx[e]] = 1;
  }
  P->pose_h = malloc(sizeof(int) * (n_poses + 1));
  P->point_h = malloc(sizeof(int) * (n_points + 1));
  P->fp_vertex = malloc(sizeof(int) * (n_poses + 1));
  P->fl_vertex = malloc(sizeof(int) * (n_points + 1));
  for (int i = 0; i < n_poses; ++i) {
    if (pa[i] && !(pose_fixed && pose_fixed[i])) { P->fp_vertex[P->n_fp] = i; P->pose_h[i] = P->n_fp++; }
    else P->pose_h[i] = -1;
  }
  for (int j = 0; j < n_points; ++j) {
    if (la[j] && !(point_fixed && point_fixed[j])) { P->fl_vertex[P->n_fl] = j; P->point_h[j] = P->n_fl++; }
    else P->point_h[j] = -1;
  }
  free(pa); free(la);

  P->err = calloc(2 * (size_t)(n_edges + 1), sizeof(double));
  if (P->n_fp + P->n_fl == 0) return 1; /* optimize(): "0 vertices to optimize" -> -1 (:368-371) */

  /* buildStructure (block_solver.hpp:102-256): one Hpl block per (free pose, free landmark)
   * pair, columns (= landmarks) hold their pose rows sorted (fillSparseBlockMatrixCCS). */
  {
    long long *keys = malloc(sizeof(long long) * (P->n_active + 1));
    int nk = 0;
    for (int k = 0; k < P->n_active; ++k) {
      int e = P->active_edge[k];
      int ip = P->pose_h[pose_idx[e]], il = P->point_h[point_idx[e]];
      if (ip >= 0 && il >= 0) keys[nk++] = (long long)il * (P->n_fp + 1) + ip;
    }
    qsort(keys, nk, sizeof(long long), cmp_ll);
    int nu = 0;
    for (int i = 0; i < nk; ++i) if (i == 0 || keys[i] != keys[i - 1]) keys[nu++] = keys[i];
    P->n_pairs = nu;
    P->pair_pose = malloc(sizeof(int) * (nu + 1));
    P->pair_lm = malloc(sizeof(int) * (nu + 1));
    P->lm_pair_ptr = calloc(P->n_fl + 2, sizeof(int));
    for (int i = 0; i < nu; ++i) {
      P->pair_lm[i] = (int)(keys[i] / (P->n_fp + 1));
      P->pair_pose[i] = (int)(keys[i] % (P->n_fp + 1));
      P->lm_pair_ptr[P->pair_lm[i] + 1]++;
    }
    for (int l = 0; l < P->n_fl; ++l) P->lm_pair_ptr[l + 1] += P->lm_pair_ptr[l];
    P->edge_pair = malloc(sizeof(int) * (P->n_active + 1));
    for (int k = 0; k < P->n_active; ++k) {
      int e = P->active_edge[k];
      int ip = P->pose_h[pose_idx[e]], il = P->point_h[point_idx[e]];
      P->edge_pair[k] = -1;
      if (ip >= 0 && il >= 0) {
        int lo = P->lm_pair_ptr[il], hi = P->lm_pair_ptr[il + 1];
        for (int a = lo; a < hi; ++a) if (P->pair_pose[a] == ip) { P->edge_pair[k] = a; break; }
      }
    }
    free(keys);
  }
  {
    size_t n = 6 * (size_t)P->n_fp;
    P->Hpp = malloc(sizeof(double) * 36 * (P->n_fp + 1));
    P->Hll = malloc(sizeof(double) * 9 * (P->n_fl + 1));
    P->Dinv = malloc(sizeof(double) * 9 * (P->n_fl + 1));
    P->Hpl = malloc(sizeof(double) * 18 * (P->n_pairs + 1));
    P->b = malloc(sizeof(double) * (n + 3 * P->n_fl + 1));
    P->x = calloc(n + 3 * P->n_fl + 1, sizeof(double));
    P->S = malloc(sizeof(double) * (n * n + 1));
    P->first = malloc(sizeof(int) * (n + 1));
    P->bschur = malloc(sizeof(double) * (n + 1));
    P->coeff = malloc(sizeof(double) * (n + 1));
  }
  return 0;
}

static void prob_free(Prob *P) {
  free(P->cam0); free(P->pose); free(P->point); free(P->pose_bak); free(P->point_bak);
  free(P->pose_h); free(P->point_h); free(P->active_edge); free(P->fp_vertex); free(P->fl_vertex);
  free(P->edge_pair); free(P->pair_pose); free(P->pair_lm); free(P->lm_pair_ptr);
  free(P->Hpp); free(P->Hll); free(P->Hpl); free(P->b); free(P->x); free(P->err); free(P->Dinv);
  free(P->S); free(P->first); free(P->bschur); free(P->coeff);
}

/* The reduced pose system at the initial estimate for one landmark shard (owner[point] == rank;
 * owner == NULL: all landmarks): S_out is (6 n_fp)^2 row-major with the upper block triangle
 * filled (as BlockSolver keeps it), b_out has 6 n_fp entries.  Summing the outputs of all ranks
 * gives the full system — the identity behind the multi-GPU all-reduce. */
int ssba_oracle_reduced_system(const double K[9], int32_t n_cams, const double *ext_qt,
                               int32_t n_poses, const double *poses_qt, const uint8_t *pose_fixed,
                               int32_t n_points, const double *points, const uint8_t *point_fixed,
                               int32_t n_edges, const int32_t *pose_idx, const int32_t *point_idx,
                               const uint8_t *cam_idx, const double *uv, double huber_delta,
                               double lambda, const int32_t *owner, int32_t rank,
                               int32_t max_free_poses, double *S_out, double *b_out, int32_t *n_free_poses) {
  (void)n_cams;
  Prob Ps, *P = &Ps;
  int empty = prob_setup(P, K, ext_qt, n_poses, poses_qt, pose_fixed, n_points, points, point_fixed,
                         n_edges, pose_idx, point_idx, cam_idx, uv, huber_delta, SSBA_JACOBIAN_ANALYTIC);
  if (n_free_poses) *n_free_poses = P->n_fp;
  int rc = 0;
  if (empty || P->n_fp > max_free_poses) { rc = 1; }
  else {
    P->owner = owner; P->rank = rank;
    compute_active_errors(P);
    build_system(P);
    assemble_reduced_system(P, lambda);
    size_t n = 6 * (size_t)P->n_fp;
    memcpy(S_out, P->S, sizeof(double) * n * n);
    memcpy(b_out, P->bschur, sizeof(double) * n);
  }
  prob_free(P);
  return rc;
}

/* _tau of OptimizationAlgorithmLevenberg (levenberg.cpp:44-51, default 1e-5); a test knob: a tiny
 * tau starts LM as Gauss-Newton, which makes the first trials overshoot and get rejected. */
static double g_tau = 1e-5;
void ssba_oracle_set_tau(double tau) { g_tau = tau; }

int ssba_oracle_optimize(const double K[9], int32_t n_cams, const double *ext_qt,
                         int32_t n_poses, const double *poses_qt, const uint8_t *pose_fixed,
                         int32_t n_points, const double *points, const uint8_t *point_fixed,
                         int32_t n_edges, const int32_t *pose_idx, const int32_t *point_idx,
                         const uint8_t *cam_idx, const double *uv, double huber_delta,
                         int32_t max_iters, int32_t jacobian_mode,
                         double *poses_out, double *points_out, double *edge_err_out,
                         ssba_report *report) {
  (void)n_cams;
  Prob Ps, *P = &Ps;
  if (report) memset(report, 0, sizeof(*report));
  double t_start = now_s();
  int rc = 0, iterations = -1;
  if (prob_setup(P, K, ext_qt, n_poses, poses_qt, pose_fixed, n_points, points, point_fixed, n_edges,
                 pose_idx, point_idx, cam_idx, uv, huber_delta, jacobian_mode)) {
    if (report) { report->iterations = -1; report->last_result = SSBA_SOLVER_FAIL; }
    goto finish;
  }

  compute_active_errors(P);
  if (report) report->chi2_initial = active_chi2(P, 1);

  /* optimize() outer loop (sparse_optimizer.cpp:386-426) around
   * OptimizationAlgorithmLevenberg::solve (levenberg.cpp:58-150) */
  const double tau = g_tau, good_upper = 2. / 3., good_lower = 1. / 3.;
  const int max_trials = 10;
  double lambda = -1., ni = 2.;
  int result = SSBA_SOLVER_OK;
  iterations = 0;
  for (int it = 0; it < max_iters && result == SSBA_SOLVER_OK; ++it) {
    compute_active_errors(P);
    double currentChi = active_chi2(P, 1), tempChi = currentChi;
    build_system(P);
    if (it == 0) { lambda = lambda_init(P, tau); ni = 2; }
    double rho = 0;
    int qmax = 0;
    const int nx = 6 * P->n_fp + 3 * P->n_fl;
    do {
      memcpy(P->pose_bak, P->pose, sizeof(double) * 7 * n_poses);   /* push() */
      memcpy(P->point_bak, P->point, sizeof(double) * 3 * n_points);
      int ok2 = solve_damped(P, lambda);
      if (!ok2 && report) report->cholesky_failures++;
      apply_update(P);
      compute_active_errors(P);
      tempChi = active_chi2(P, 1);
      if (!ok2) tempChi = DBL_MAX;
      rho = currentChi - tempChi;
      double scale = 0; /* computeScale (:168-175) */
      for (int j = 0; j < nx; ++j) scale += P->x[j] * (lambda * P->x[j] + P->b[j]);
      scale += 1e-3;
      rho /= scale;
      if (rho > 0 && isfinite(tempChi)) {
        double alpha = 1. - pow((2 * rho - 1), 3);
        alpha = alpha < good_upper ? alpha : good_upper;
        double scaleFactor = good_lower > alpha ? good_lower : alpha;
        lambda *= scaleFactor;
        ni = 2;
        currentChi = tempChi;
      } else {
        lambda *= ni;
        ni *= 2;
        memcpy(P->pose, P->pose_bak, sizeof(double) * 7 * n_poses);  /* pop() */
        memcpy(P->point, P->point_bak, sizeof(double) * 3 * n_points);
        if (!isfinite(lambda)) break;
      }
      qmax++;
    } while (rho < 0 && qmax < max_trials);
    if (qmax == max_trials || rho == 0 || !isfinite(lambda)) result = SSBA_SOLVER_TERMINATE;
    ++iterations;
    if (report && report->n_records < SSBA_MAX_ITER_RECORDS) {
      ssba_iter_record *r = &report->iters[report->n_records++];
      compute_active_errors(P);
      r->chi2 = active_chi2(P, 1);
      r->lambda = lambda; r->trials = qmax; r->result = result;
    }
  }
  if (report) { report->iterations = iterations; report->last_result = result; report->lambda = lambda; }

finish:
  if (iterations >= 0) {
    compute_active_errors(P);
    if (report) { report->chi2_robust = active_chi2(P, 1); report->chi2_plain = active_chi2(P, 0); }
  }
  if (report) report->seconds_total = now_s() - t_start;
  if (poses_out) memcpy(poses_out, P->pose, sizeof(double) * 7 * n_poses);
  if (points_out) memcpy(points_out, P->point, sizeof(double) * 3 * n_points);
  if (edge_err_out) memcpy(edge_err_out, P->err, sizeof(double) * 2 * n_edges);
  prob_free(P);
  return rc;
}
