// ssba_pose_only.cu — batched pose-only Levenberg-Marquardt: FrontEnd::EstimateCurrentPose
// (src/ssvio/frontend.cpp:184-260) for many frames at once, one warp per frame, the whole
// optimisation (rounds x iterations x trials, outlier re-classification) in ONE launch.
//
// Reference semantics reproduced here (g2o/ = thirdparty/g2o/g2o/):
//   EdgeProjectionPoseOnly::computeError / linearizeOplus   include/ssvio/g2otypes.hpp:78-101
//       (analytic Jacobian AS SHIPPED, Zinv = 1 / (Z + 1e-18))
//   BaseUnaryEdge::constructQuadraticForm                    g2o/core/base_unary_edge.hpp:49-79
//   RobustKernelHuber with its default delta = 1             g2o/core/robust_kernel_impl.cpp:65-78
//   OptimizationAlgorithmLevenberg::solve + computeScale     g2o/core/optimization_algorithm_levenberg.cpp:58-175
//   SparseOptimizer::optimize                                g2o/core/sparse_optimizer.cpp:366-431
//   LinearSolverDense (a failed factorisation rejects)       g2o/solvers/dense/linear_solver_dense.h:56-118
//   the round loop, chi2() > threshold, setLevel, kernel removal at round rounds - 2
//                                                            src/ssvio/frontend.cpp:235-270
//   LoopClosing::OptimizeCurrentPose (src/ssvio/loopclosing.cpp:245-351): the same loop after one more
//       initializeOptimization(); optimize(10) with every edge and the kernel on (:302-303) = pre_rounds 1
// The edges' _error members live in `err`: like in g2o they hold the errors of the LAST evaluated
// trial (a rejected trial does not restore them), and that is what the outlier test reads.
#include <cfloat>

#include "ssba_device.hpp"

namespace ssba {

namespace {

__device__ __forceinline__ double warp_allsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// e = z - (K (T p)) / depth; P = T p is returned for the Jacobian
__device__ __forceinline__ void po_error(const double *K, const double *T, const double *p, double u, double v,
                                         double &e0, double &e1, double *P) {
  se3_act(T, p[0], p[1], p[2], P[0], P[1], P[2]);
  const double n0 = K[0] * P[0] + K[1] * P[1] + K[2] * P[2];
  const double n1 = K[3] * P[0] + K[4] * P[1] + K[5] * P[2];
  const double dn = K[6] * P[0] + K[7] * P[1] + K[8] * P[2];
  e0 = u - n0 / dn;
  e1 = v - n1 / dn;
}

// 6x6 Cholesky solve (H + lambda I) x = b, H given by its upper triangle (21); false on a pivot <= 0
__device__ __forceinline__ bool po_solve(const double *Hu, double lambda, const double *b, double *x) {
  double a[36];
  int k = 0;
#pragma unroll
  for (int r = 0; r < 6; ++r)
#pragma unroll
    for (int c = r; c < 6; ++c) { a[6 * c + r] = Hu[k]; a[6 * r + c] = Hu[k]; ++k; }
#pragma unroll
  for (int d = 0; d < 6; ++d) a[7 * d] += lambda;
  bool ok = true;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    double dj = a[7 * j];
    if (!(dj > 0.0)) { ok = false; dj = 1.0; }
    const double inv = 1.0 / sqrt(dj);
    a[7 * j] = sqrt(dj);
#pragma unroll
    for (int i = j + 1; i < 6; ++i) a[6 * i + j] *= inv;
#pragma unroll
    for (int i = j + 1; i < 6; ++i)
#pragma unroll
      for (int c = j + 1; c <= i; ++c) a[6 * i + c] -= a[6 * i + j] * a[6 * c + j];
  }
  double y[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    double s = b[i];
#pragma unroll
    for (int c = 0; c < i; ++c) s -= a[6 * i + c] * y[c];
    y[i] = s / a[7 * i];
  }
#pragma unroll
  for (int i = 5; i >= 0; --i) {
    double s = y[i];
#pragma unroll
    for (int c = i + 1; c < 6; ++c) s -= a[6 * c + i] * x[c];
    x[i] = s / a[7 * i];
  }
  return ok;
}

struct PoseOnlyArgs {
  double K[9];
  int n_frames, rounds, iters, max_trials, pre_rounds;
  double chi2_threshold, tau, good_lower, good_upper, user_lambda;
  const int32_t *feat_ptr;
  const double *poses_in, *xyz, *uv;
  double *err;        // N x 2 scratch: the edges' _error
  uint8_t *outlier;   // N (in/out scratch, final result)
  double *poses_out, *chi2_out;
  int32_t *n_inliers_out;
};

constexpr int kPoWarps = 4;

__global__ void __launch_bounds__(32 * kPoWarps) k_pose_only(const PoseOnlyArgs A) {
  const int lane = threadIdx.x & 31;
  const int f = blockIdx.x * kPoWarps + (threadIdx.x >> 5);
  if (f >= A.n_frames) return;  // whole warp
  const int e0 = A.feat_ptr[f], n = A.feat_ptr[f + 1] - e0;
  double T[7];
#pragma unroll
  for (int i = 0; i < 7; ++i) T[i] = A.poses_in[7 * f + i];
  const double *xyz = A.xyz + 3 * (size_t)e0, *uv = A.uv + 2 * (size_t)e0;
  double *err = A.err + 2 * (size_t)e0;
  uint8_t *outl = A.outlier + e0;  // outlier flag == level 1 (frontend.cpp:252-260)
  for (int i = lane; i < n; i += 32) outl[i] = 0;
  __syncwarp();
  bool robust = true;
  double chi_last = 0.0;
  int cnt_out = 0;
  // errors of the active edges at pose Tq -> err, returns the (robust) chi2 of the active set
  auto eval = [&](const double *Tq) {
    double s = 0.0;
    for (int i = lane; i < n; i += 32) {
      if (outl[i]) continue;
      double P[3], a, b;
      po_error(A.K, Tq, xyz + 3 * i, uv[2 * i], uv[2 * i + 1], a, b, P);
      err[2 * i] = a; err[2 * i + 1] = b;
      double r0 = a * a + b * b, r1;
      if (robust) huber(r0, 1.0, r0, r1);
      s += r0;
    }
    return warp_allsum(s);
  };
  for (int rnd = -A.pre_rounds; rnd < A.rounds; ++rnd) {
    int n_act = 0;
    for (int i = lane; i < n; i += 32) n_act += outl[i] ? 0 : 1;
    n_act = __reduce_add_sync(0xffffffffu, n_act);
    if (n_act > 0) {
      double lambda = 0.0, ni = 2.0;
      if (A.iters <= 0) chi_last = eval(T);  // no iteration will evaluate the active edges: the test below reads err
      for (int it = 0; it < A.iters; ++it) {
        const double cur0 = eval(T);
        // buildSystem: H (upper triangle) and b over the active edges
        double acc[27];
#pragma unroll
        for (int i = 0; i < 27; ++i) acc[i] = 0.0;
        const double fx = A.K[0], fy = A.K[4];
        for (int i = lane; i < n; i += 32) {
          if (outl[i]) continue;
          double P[3];
          se3_act(T, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], P[0], P[1], P[2]);
          const double X = P[0], Y = P[1], Z = P[2];
          const double zi = 1.0 / (Z + 1e-18), zi2 = zi * zi;
          const double J0[6] = {-fx * zi, 0.0, fx * X * zi2, fx * X * Y * zi2, -fx - fx * X * X * zi2, fx * Y * zi};
          const double J1[6] = {0.0, -fy * zi, fy * Y * zi2, fy + fy * Y * Y * zi2, -fy * X * Y * zi2, -fy * X * zi};
          const double a = err[2 * i], b = err[2 * i + 1];
          double w = 1.0, r0 = a * a + b * b;
          if (robust) huber(r0, 1.0, r0, w);
          int k = 6;
#pragma unroll
          for (int r = 0; r < 6; ++r) {
            acc[r] -= w * (J0[r] * a + J1[r] * b);
#pragma unroll
            for (int c = r; c < 6; ++c) acc[k++] += w * (J0[r] * J0[c] + J1[r] * J1[c]);
          }
        }
#pragma unroll
        for (int i = 0; i < 27; ++i) acc[i] = warp_allsum(acc[i]);
        const double *bvec = acc, *Hu = acc + 6;
        if (it == 0) {  // computeLambdaInit (levenberg.cpp:152-166)
          double m = 0.0;
#pragma unroll
          for (int d = 0; d < 6; ++d) m = fmax(m, fabs(Hu[d * 6 - d * (d - 1) / 2]));
          lambda = A.user_lambda > 0 ? A.user_lambda : A.tau * m; ni = 2.0;
        }
        double cur = cur0, rho = 0.0;
        int qmax = 0;
        bool stop = false;
        do {
          double x[6], Tn[7];
          const bool ok = po_solve(Hu, lambda, bvec, x);
          if (ok) pose_oplus(T, x, Tn);
          else {
#pragma unroll
            for (int i = 0; i < 7; ++i) Tn[i] = T[i];
          }
          double tmp = eval(Tn);
          chi_last = tmp;
          if (!ok) tmp = DBL_MAX;
          double scale = 1e-3;
#pragma unroll
          for (int i = 0; i < 6; ++i) scale += x[i] * (lambda * x[i] + bvec[i]);
          rho = (cur - tmp) / scale;
          if (!ok) rho = -1.0;
          if (rho > 0 && isfinite(tmp)) {
            const double t = 2 * rho - 1;
            const double alpha = fmin(1.0 - t * t * t, A.good_upper);
            lambda *= fmax(A.good_lower, alpha);
            ni = 2.0;
            cur = tmp;
#pragma unroll
            for (int i = 0; i < 7; ++i) T[i] = Tn[i];
          } else {
            lambda *= ni;
            ni *= 2;
            if (!isfinite(lambda)) { stop = true; break; }
          }
          ++qmax;
        } while (rho < 0 && qmax < A.max_trials);
        if (qmax == A.max_trials || rho == 0 || stop) break;  // Terminate
      }
    }
    if (rnd < 0) continue;  // loopclosing.cpp:302-303: an optimize() without classification
    // re-classification (frontend.cpp:243-262)
    cnt_out = 0;
    for (int i = lane; i < n; i += 32) {
      double a, b;
      if (outl[i]) {
        double P[3];
        po_error(A.K, T, xyz + 3 * i, uv[2 * i], uv[2 * i + 1], a, b, P);
        err[2 * i] = a; err[2 * i + 1] = b;
      } else {
        a = err[2 * i]; b = err[2 * i + 1];
      }
      const bool o = a * a + b * b > A.chi2_threshold;
      outl[i] = o ? 1 : 0;
      cnt_out += o ? 1 : 0;
    }
    cnt_out = __reduce_add_sync(0xffffffffu, cnt_out);
    __syncwarp();
    if (rnd == A.rounds - 2) robust = false;
  }
  if (lane < 7) A.poses_out[7 * f + lane] = T[lane];
  if (lane == 0) { A.n_inliers_out[f] = n - cnt_out; A.chi2_out[f] = chi_last; }
}

}  // namespace

void launch_pose_only(const double K[9], int n_frames, int rounds, int iters, int pre_rounds, int max_trials, double chi2_threshold,
                      double tau, double good_lower, double good_upper, double user_lambda, const int32_t *feat_ptr, const double *poses_in,
                      const double *xyz, const double *uv, double *err, uint8_t *outlier, double *poses_out,
                      double *chi2_out, int32_t *n_inliers_out, cudaStream_t st) {
  if (n_frames <= 0) return;
  PoseOnlyArgs A;
  for (int i = 0; i < 9; ++i) A.K[i] = K[i];
  A.n_frames = n_frames; A.rounds = rounds; A.iters = iters; A.max_trials = max_trials; A.pre_rounds = pre_rounds > 0 ? pre_rounds : 0;
  A.chi2_threshold = chi2_threshold; A.tau = tau; A.good_lower = good_lower; A.good_upper = good_upper; A.user_lambda = user_lambda;
  A.feat_ptr = feat_ptr; A.poses_in = poses_in; A.xyz = xyz; A.uv = uv; A.err = err; A.outlier = outlier;
  A.poses_out = poses_out; A.chi2_out = chi2_out; A.n_inliers_out = n_inliers_out;
  k_pose_only<<<(n_frames + kPoWarps - 1) / kPoWarps, 32 * kPoWarps, 0, st>>>(A);
}

}  // namespace ssba
