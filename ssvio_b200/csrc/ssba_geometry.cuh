// ssba_geometry.cuh — fp64 device arithmetic of the local-BA hot path.
//
// Re-derivation (not a translation) of the fixed-size algebra the reference gets from Sophus and
// Eigen; each function names the reference semantics it has to reproduce.  Paths relative to the
// ssvio tree, g2o/ = thirdparty/g2o/g2o/, sophus/ = thirdparty/sophus/sophus/.
#pragma once

#include <cuda_runtime.h>
#include <math.h>

#define SSBA_HD __host__ __device__ __forceinline__

namespace ssba {

struct Cameras {
  double K[9];        // row-major intrinsics (g2otypes.hpp:118-121)
  double ext[8][7];   // camera <- body, qx qy qz qw tx ty tz (backend.cpp:105-107)
  int n;
};

// Sophus SO3 point action, quaternion sandwich p + w*2(v x p) + v x 2(v x p) (sophus/so3.hpp:352-360)
SSBA_HD void quat_rotate(const double *q, double px, double py, double pz, double &ox, double &oy,
                         double &oz) {
  double ux = q[1] * pz - q[2] * py, uy = q[2] * px - q[0] * pz, uz = q[0] * py - q[1] * px;
  ux += ux; uy += uy; uz += uz;
  ox = px + q[3] * ux + (q[1] * uz - q[2] * uy);
  oy = py + q[3] * uy + (q[2] * ux - q[0] * uz);
  oz = pz + q[3] * uz + (q[0] * uy - q[1] * ux);
}

// SE3 point action (sophus/se3.hpp:325-328)
SSBA_HD void se3_act(const double *T, double px, double py, double pz, double &ox, double &oy,
                     double &oz) {
  quat_rotate(T, px, py, pz, ox, oy, oz);
  ox += T[4]; oy += T[5]; oz += T[6];
}

// rotation matrix of a unit quaternion, row-major (Eigen toRotationMatrix)
SSBA_HD void quat_to_matrix(const double *q, double *R) {
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w;
  const double txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
  R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}

// VertexPose::oplusImpl (g2otypes.hpp:36-41): T <- SE3::exp(d) * T, with
// SE3::exp (sophus/se3.hpp:763-784), SO3::expAndTheta incl. its theta^2 < 1e-20 Taylor branch
// (sophus/so3.hpp:593-622), and the quaternion re-normalisation every SO3 product performs
// (sophus/so3.hpp:322-334,498-503).
SSBA_HD void pose_oplus(const double *T, const double *d, double *out) {
  const double eps = 1e-10;
  const double wx = d[3], wy = d[4], wz = d[5];
  const double theta_sq = wx * wx + wy * wy + wz * wz;
  double theta, imag, real;
  if (theta_sq < eps * eps) {
    const double t4 = theta_sq * theta_sq;
    theta = 0.0;
    imag = 0.5 - (1.0 / 48.0) * theta_sq + (1.0 / 3840.0) * t4;
    real = 1.0 - (1.0 / 8.0) * theta_sq + (1.0 / 384.0) * t4;
  } else {
    theta = sqrt(theta_sq);
    const double half = 0.5 * theta;
    imag = sin(half) / theta;
    real = cos(half);
  }
  double e[7];
  e[0] = imag * wx; e[1] = imag * wy; e[2] = imag * wz; e[3] = real;
  // translation of exp(d): V * upsilon
  double V[9];
  if (theta < eps) {
    quat_to_matrix(e, V);
  } else {
    const double th2 = theta * theta;
    const double c1 = (1.0 - cos(theta)) / th2;
    const double c2 = (theta - sin(theta)) / (th2 * theta);
    // Omega = hat(w), Omega^2 = w w^T - |w|^2 I
    V[0] = 1.0 + c2 * (-wy * wy - wz * wz); V[1] = -c1 * wz + c2 * (wx * wy); V[2] = c1 * wy + c2 * (wx * wz);
    V[3] = c1 * wz + c2 * (wx * wy); V[4] = 1.0 + c2 * (-wx * wx - wz * wz); V[5] = -c1 * wx + c2 * (wy * wz);
    V[6] = -c1 * wy + c2 * (wx * wz); V[7] = c1 * wx + c2 * (wy * wz); V[8] = 1.0 + c2 * (-wx * wx - wy * wy);
  }
  e[4] = V[0] * d[0] + V[1] * d[1] + V[2] * d[2];
  e[5] = V[3] * d[0] + V[4] * d[1] + V[5] * d[2];
  e[6] = V[6] * d[0] + V[7] * d[1] + V[8] * d[2];
  // SE3 product e * T (sophus/se3.hpp:309-314)
  const double ax = e[0], ay = e[1], az = e[2], aw = e[3];
  const double bx = T[0], by = T[1], bz = T[2], bw = T[3];
  const double w = aw * bw - ax * bx - ay * by - az * bz;
  const double x = aw * bx + ax * bw + ay * bz - az * by;
  const double y = aw * by + ay * bw + az * bx - ax * bz;
  const double z = aw * bz + az * bw + ax * by - ay * bx;
  const double len = sqrt(x * x + y * y + z * z + w * w);
  out[0] = x / len; out[1] = y / len; out[2] = z / len; out[3] = w / len;
  double rx, ry, rz;
  quat_rotate(e, T[4], T[5], T[6], rx, ry, rz);
  out[4] = e[4] + rx; out[5] = e[5] + ry; out[6] = e[6] + rz;
}

// EdgeProjection::computeError (g2otypes.hpp:123-131): e = z - (K (ext (T p))) / depth.
// Returns the body-frame and camera-frame points as by-products for the Jacobians.
SSBA_HD void edge_error(const double *K, const double *ext, const double *T, const double *p,
                        double u, double v, double &e0, double &e1) {
  double bx, by, bz, cx, cy, cz;
  se3_act(T, p[0], p[1], p[2], bx, by, bz);
  se3_act(ext, bx, by, bz, cx, cy, cz);
  const double n0 = K[0] * cx + K[1] * cy + K[2] * cz;
  const double n1 = K[3] * cx + K[4] * cy + K[5] * cz;
  const double dn = K[6] * cx + K[7] * cy + K[8] * cz;
  e0 = u - n0 / dn;
  e1 = v - n1 / dn;
}

// RobustKernelHuber::robustify (g2o/core/robust_kernel_impl.cpp:65-78): rho0 and rho1 of the
// squared error e2; delta <= 0 means "no robust kernel" (base_binary_edge.hpp:82).
SSBA_HD void huber(double e2, double delta, double &rho0, double &rho1) {
  const double dsqr = delta * delta;
  if (delta <= 0.0 || e2 <= dsqr) {
    rho0 = e2; rho1 = 1.0;
  } else {
    const double s = sqrt(e2);
    rho0 = 2 * s * delta - dsqr;
    rho1 = delta / s;
  }
}

// Residual and closed-form Jacobians of one edge:
//   J_xi = Jpi R_ext [ I , -[T p]x ]   (2x6, tangent order upsilon, omega; left-multiplicative)
//   J_p  = Jpi R_ext R                 (2x3)
// Jpi = d(-proj)/dP at the camera-frame point, written for a general K.
// Re (row-major 3x3) is the precomputed rotation matrix of the extrinsic.
SSBA_HD void edge_linearize_analytic(const double *K, const double *ext, const double *Re,
                                     const double *T, const double *p, double u, double v,
                                     double &e0, double &e1, double *Jx /*12*/, double *Jp /*6*/) {
  double bx, by, bz, cx, cy, cz;
  se3_act(T, p[0], p[1], p[2], bx, by, bz);
  se3_act(ext, bx, by, bz, cx, cy, cz);
  const double n0 = K[0] * cx + K[1] * cy + K[2] * cz;
  const double n1 = K[3] * cx + K[4] * cy + K[5] * cz;
  const double dn = K[6] * cx + K[7] * cy + K[8] * cz;
  const double id = 1.0 / dn;
  e0 = u - n0 / dn;
  e1 = v - n1 / dn;
  const double id2 = id * id;
  double Jpi[6];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    Jpi[c] = -(K[c] * id - n0 * K[6 + c] * id2);
    Jpi[3 + c] = -(K[3 + c] * id - n1 * K[6 + c] * id2);
  }
  double R[9];
  quat_to_matrix(T, R);
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const double j0 = Jpi[3 * r] * Re[0] + Jpi[3 * r + 1] * Re[3] + Jpi[3 * r + 2] * Re[6];
    const double j1 = Jpi[3 * r] * Re[1] + Jpi[3 * r + 1] * Re[4] + Jpi[3 * r + 2] * Re[7];
    const double j2 = Jpi[3 * r] * Re[2] + Jpi[3 * r + 1] * Re[5] + Jpi[3 * r + 2] * Re[8];
    Jx[6 * r + 0] = j0; Jx[6 * r + 1] = j1; Jx[6 * r + 2] = j2;
    Jx[6 * r + 3] = -(j1 * bz - j2 * by);
    Jx[6 * r + 4] = -(j2 * bx - j0 * bz);
    Jx[6 * r + 5] = -(j0 * by - j1 * bx);
    Jp[3 * r + 0] = j0 * R[0] + j1 * R[3] + j2 * R[6];
    Jp[3 * r + 1] = j0 * R[1] + j1 * R[4] + j2 * R[7];
    Jp[3 * r + 2] = j0 * R[2] + j1 * R[5] + j2 * R[8];
  }
}

// Central-difference Jacobians, delta = 1e-9: the reference AS SHIPPED
// (g2o/core/base_binary_edge.hpp:144-212; the analytic override is commented out,
// g2otypes.hpp:133-153).  18 extra error evaluations per edge.
SSBA_HD void edge_linearize_numeric(const double *K, const double *ext, const double *T,
                                    const double *p, double u, double v, double &e0, double &e1,
                                    double *Jx, double *Jp) {
  const double delta = 1e-9;
  const double scalar = 1 / (2 * delta);
  edge_error(K, ext, T, p, u, v, e0, e1);
#pragma unroll 1
  for (int d = 0; d < 6; ++d) {
    double add[6] = {0, 0, 0, 0, 0, 0}, T1[7], a0, a1, b0, b1;
    add[d] = delta;
    pose_oplus(T, add, T1);
    edge_error(K, ext, T1, p, u, v, a0, a1);
    add[d] = -delta;
    pose_oplus(T, add, T1);
    edge_error(K, ext, T1, p, u, v, b0, b1);
    Jx[d] = scalar * (a0 - b0);
    Jx[6 + d] = scalar * (a1 - b1);
  }
#pragma unroll 1
  for (int d = 0; d < 3; ++d) {
    double p1[3] = {p[0], p[1], p[2]}, a0, a1, b0, b1;
    p1[d] = p[d] + delta;
    edge_error(K, ext, T, p1, u, v, a0, a1);
    p1[d] = p[d] + (-delta);
    edge_error(K, ext, T, p1, u, v, b0, b1);
    Jp[d] = scalar * (a0 - b0);
    Jp[3 + d] = scalar * (a1 - b1);
  }
}

// Inverse of a symmetric 3x3 (m = xx xy xz yy yz zz) by cofactors / determinant — the closed
// form Eigen 3.3.7 uses for Matrix3d::inverse() at g2o/core/block_solver.hpp:350.
SSBA_HD void sym3_inverse(const double *m, double *o) {
  const double a = m[0], b = m[1], c = m[2], d = m[3], e = m[4], f = m[5];
  const double c00 = d * f - e * e;
  const double c01 = c * e - b * f;
  const double c02 = b * e - c * d;
  const double det = a * c00 + b * c01 + c * c02;
  const double id = 1.0 / det;
  o[0] = c00 * id;
  o[1] = c01 * id;
  o[2] = c02 * id;
  o[3] = (a * f - c * c) * id;
  o[4] = (b * c - a * e) * id;
  o[5] = (a * d - b * b) * id;
}

// ---- SE(3) group operations of the pose-graph edge (EdgePoseGraph, g2otypes.hpp:169-176)

// SE3 product a * b with the quaternion re-normalised like every Sophus SO3 product
// (sophus/se3.hpp:309-314, so3.hpp:322-334,498-503)
SSBA_HD void se3_mul(const double *a, const double *b, double *o) {
  const double ax = a[0], ay = a[1], az = a[2], aw = a[3];
  const double bx = b[0], by = b[1], bz = b[2], bw = b[3];
  const double w = aw * bw - ax * bx - ay * by - az * bz;
  const double x = aw * bx + ax * bw + ay * bz - az * by;
  const double y = aw * by + ay * bw + az * bx - ax * bz;
  const double z = aw * bz + az * bw + ax * by - ay * bx;
  const double len = sqrt(x * x + y * y + z * z + w * w);
  double rx, ry, rz;
  quat_rotate(a, b[4], b[5], b[6], rx, ry, rz);
  o[0] = x / len; o[1] = y / len; o[2] = z / len; o[3] = w / len;
  o[4] = a[4] + rx; o[5] = a[5] + ry; o[6] = a[6] + rz;
}

// SE3 inverse (sophus/se3.hpp:214-217): (q*, -q* t)
SSBA_HD void se3_inverse(const double *a, double *o) {
  o[0] = -a[0]; o[1] = -a[1]; o[2] = -a[2]; o[3] = a[3];
  double rx, ry, rz;
  quat_rotate(o, a[4], a[5], a[6], rx, ry, rz);
  o[4] = -rx; o[5] = -ry; o[6] = -rz;
}

// SE3::log (sophus/se3.hpp:223-256) with SO3::logAndTheta (sophus/so3.hpp:245-286): tangent (upsilon, omega)
SSBA_HD void se3_log(const double *T, double *d) {
  const double eps = 1e-10;
  const double sq = T[0] * T[0] + T[1] * T[1] + T[2] * T[2], w = T[3];
  double f, theta;
  if (sq < eps * eps) {
    f = 2.0 / w - (2.0 / 3.0) * sq / (w * w * w);
    theta = 2.0 * sq / w;
  } else {
    const double n = sqrt(sq);
    if (fabs(w) < eps) f = (w > 0.0 ? 3.14159265358979323846 : -3.14159265358979323846) / n;
    else f = 2.0 * atan(n / w) / n;
    theta = f * n;
  }
  const double ox = f * T[0], oy = f * T[1], oz = f * T[2];
  const double tx = T[4], ty = T[5], tz = T[6];
  // V^-1 t = t - 0.5 (omega x t) + c (omega x (omega x t))
  const double ax = oy * tz - oz * ty, ay = oz * tx - ox * tz, az = ox * ty - oy * tx;
  const double bx = oy * az - oz * ay, by = oz * ax - ox * az, bz = ox * ay - oy * ax;
  double c;
  if (fabs(theta) < eps) c = 1.0 / 12.0;
  else { const double h = 0.5 * theta; c = (1.0 - theta * cos(h) / (2.0 * sin(h))) / (theta * theta); }
  d[0] = tx - 0.5 * ax + c * bx; d[1] = ty - 0.5 * ay + c * by; d[2] = tz - 0.5 * az + c * bz;
  d[3] = ox; d[4] = oy; d[5] = oz;
}

// EdgePoseGraph::computeError (g2otypes.hpp:169-176): e = log(Minv * T0 * T1^-1); Minv = measurement^-1
SSBA_HD void pose_graph_error(const double *Minv, const double *T0, const double *T1, double *e) {
  double a[7], b[7], c[7];
  se3_mul(Minv, T0, a);
  se3_inverse(T1, b);
  se3_mul(a, b, c);
  se3_log(c, e);
}

}  // namespace ssba
