// ssba_kernels.cu — sm_100a kernels of the local-BA hot path (first correct version).
//
// One LM trial ("slot") is the stream-ordered sequence
//   [linearize + hpp]  prepare_system  schur  (all-reduce)  reduced_solve  update  (all-reduce)  control
// Every kernel reads the device-resident Control block first and returns at once when the
// optimisation is finished or its phase is not needed, so the host can enqueue slots ahead
// without synchronising per trial (the accept/reject decision of levenberg.cpp:128-143 is
// taken on the device by k_control).
//
// Reference semantics per kernel are cited at each kernel; g2o/ = thirdparty/g2o/g2o/.
#include <cfloat>

#include "ssba_device.hpp"

namespace ssba {

namespace {

constexpr int kLinThreads = 128;
constexpr int kHppThreads = 128;
constexpr int kSolveThreads = 256;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
  return v;
}

// deterministic block reductions: warp tree, then warp 0 lane 0 folds the warps in order
template <int NT>
__device__ __forceinline__ double block_sum(double v, double *sm /* NT/32 */) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = 0.0;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) r += sm[w];
  }
  __syncthreads();
  return r;  // valid on thread 0
}
template <int NT>
__device__ __forceinline__ double block_max(double v, double *sm) {
  v = warp_max(v);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = 0.0;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) r = fmax(r, sm[w]);
  }
  __syncthreads();
  return r;
}

struct EdgeTerms {
  double e0, e1;      // error
  double we0, we1;    // Omega e
  double chi;         // e^T Omega e   (base_edge.h:79-82)
  double rho0, w;     // Huber rho, rho'
  double o00, o01, o11;
};

__device__ __forceinline__ void load_edge_weighting(const DeviceProblem &P, const double *info,
                                                    const double *delta, int e, EdgeTerms &t) {
  t.o00 = 1.0; t.o01 = 0.0; t.o11 = 1.0;
  if (info) { t.o00 = info[3 * e]; t.o01 = info[3 * e + 1]; t.o11 = info[3 * e + 2]; }
  t.we0 = t.o00 * t.e0 + t.o01 * t.e1;
  t.we1 = t.o01 * t.e0 + t.o11 * t.e1;
  t.chi = t.e0 * t.we0 + t.e1 * t.we1;
  const double d = delta ? delta[e] : P.delta_all;
  huber(t.chi, d, t.rho0, t.w);
}

__device__ __forceinline__ void linearize_edge(const DeviceProblem &P, int cam, const double *T,
                                               const double *p, double u, double v, EdgeTerms &t,
                                               double *Jx, double *Jp) {
  if (P.jacobian_mode == SSBA_JACOBIAN_NUMERIC)
    edge_linearize_numeric(P.cams.K, P.cams.ext[cam], T, p, u, v, t.e0, t.e1, Jx, Jp);
  else
    edge_linearize_analytic(P.cams.K, P.cams.ext[cam], P.ext_R[cam], T, p, u, v, t.e0, t.e1, Jx, Jp);
}

// ---------------------------------------------------------------------------------------------
// k_linearize — landmark side of BlockSolver::buildSystem (g2o/core/block_solver.hpp:462-521):
// per edge linearizeOplus + constructQuadraticForm (g2o/core/base_binary_edge.hpp:61-134) with
// the Huber weighting of robust_kernel_impl.cpp:65-78 / base_edge.h:117-123, accumulating
// Hll, b_l and the Hpl blocks W; also activeRobustChi2 of the current state
// (sparse_optimizer.cpp:102-116) and the landmark part of computeLambdaInit's max |H_jj|.
// One thread per landmark slot: Hll/b_l/W stay in registers, no atomics.
__global__ void __launch_bounds__(kLinThreads) k_linearize(const DeviceProblem P) {
  const Control *ctl = P.ctl;
  if (ctl->done || !ctl->need_linearize) return;
  __shared__ double red[kLinThreads / 32];
  const int cur = ctl->cur;
  const double *__restrict__ pose = P.pose[cur];
  const double *__restrict__ point = P.point[cur];
  const int sl = blockIdx.x * blockDim.x + threadIdx.x;
  double chi = 0.0, mx = 0.0;
  if (sl < P.n_slots) {
    const int pv = P.slot_vertex[sl];
    const bool lfree = P.slot_free[sl] != 0;
    const double p[3] = {point[3 * pv], point[3 * pv + 1], point[3 * pv + 2]};
    double H[6] = {0, 0, 0, 0, 0, 0}, b[3] = {0, 0, 0};
    const int a1 = P.slot_pair_ptr[sl + 1];
    for (int a = P.slot_pair_ptr[sl]; a < a1; ++a) {
      const int kv = P.pair_vertex[a];
      const bool wpair = lfree && P.pair_q[a] >= 0;
      double T[7];
#pragma unroll
      for (int i = 0; i < 7; ++i) T[i] = pose[7 * kv + i];
      double W[18];
#pragma unroll
      for (int i = 0; i < 18; ++i) W[i] = 0.0;
      const int e1 = P.pair_edge_ptr[a + 1];
      for (int e = P.pair_edge_ptr[a]; e < e1; ++e) {
        EdgeTerms t;
        double Jx[12], Jp[6];
        linearize_edge(P, P.e_cam[e], T, p, P.e_uv[2 * e], P.e_uv[2 * e + 1], t, Jx, Jp);
        load_edge_weighting(P, P.e_info, P.e_delta, e, t);
        chi += t.rho0;
        if (lfree) {
          // rows of (rho' Omega) J_p
          double A0[3], A1[3];
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            A0[c] = t.w * (t.o00 * Jp[c] + t.o01 * Jp[3 + c]);
            A1[c] = t.w * (t.o01 * Jp[c] + t.o11 * Jp[3 + c]);
          }
          H[0] += Jp[0] * A0[0] + Jp[3] * A1[0];
          H[1] += Jp[0] * A0[1] + Jp[3] * A1[1];
          H[2] += Jp[0] * A0[2] + Jp[3] * A1[2];
          H[3] += Jp[1] * A0[1] + Jp[4] * A1[1];
          H[4] += Jp[1] * A0[2] + Jp[4] * A1[2];
          H[5] += Jp[2] * A0[2] + Jp[5] * A1[2];
          const double r0 = -t.w * t.we0, r1 = -t.w * t.we1;  // omega_r * rho'
#pragma unroll
          for (int c = 0; c < 3; ++c) b[c] += Jp[c] * r0 + Jp[3 + c] * r1;
          if (wpair) {
#pragma unroll
            for (int r = 0; r < 6; ++r)
#pragma unroll
              for (int c = 0; c < 3; ++c) W[3 * r + c] += Jx[r] * A0[c] + Jx[6 + r] * A1[c];
          }
        }
      }
      if (wpair) {
        double2 *dst = reinterpret_cast<double2 *>(P.W + 18 * (size_t)a);
#pragma unroll
        for (int i = 0; i < 9; ++i) dst[i] = make_double2(W[2 * i], W[2 * i + 1]);
      }
    }
    if (lfree) {
#pragma unroll
      for (int i = 0; i < 6; ++i) P.Hll[6 * (size_t)sl + i] = H[i];
#pragma unroll
      for (int i = 0; i < 3; ++i) P.bl[3 * (size_t)sl + i] = b[i];
      mx = fmax(fabs(H[0]), fmax(fabs(H[3]), fabs(H[5])));
    }
  }
  const double s = block_sum<kLinThreads>(chi, red);
  const double m = block_max<kLinThreads>(mx, red);
  if (threadIdx.x == 0) { P.chi_cur_part[blockIdx.x] = s; P.maxdiag_part[blockIdx.x] = m; }
}

// ---------------------------------------------------------------------------------------------
// k_hpp — pose side of buildSystem: Hpp_ii += J_xi^T (rho' Omega) J_xi, b_i += J_xi^T(-rho' Omega e)
// (base_binary_edge.hpp:104-110).  Pose-major pass, one CTA per chunk of <= 256 edges of ONE pose:
// 27 register accumulators per thread, deterministic block reduction, no atomics.
__global__ void __launch_bounds__(kHppThreads) k_hpp(const DeviceProblem P) {
  const Control *ctl = P.ctl;
  if (ctl->done || !ctl->need_linearize) return;
  __shared__ double red[27][kHppThreads / 32];
  const int cur = ctl->cur;
  const double *__restrict__ pose = P.pose[cur];
  const double *__restrict__ point = P.point[cur];
  const int c = blockIdx.x;
  const int kv = P.chunk_vertex[c];
  double T[7];
#pragma unroll
  for (int i = 0; i < 7; ++i) T[i] = pose[7 * kv + i];
  double acc[27];
#pragma unroll
  for (int i = 0; i < 27; ++i) acc[i] = 0.0;
  const int e1 = P.chunk_edge_ptr[c + 1];
  for (int e = P.chunk_edge_ptr[c] + threadIdx.x; e < e1; e += kHppThreads) {
    const int pv = P.pm_point[e];
    const double p[3] = {point[3 * pv], point[3 * pv + 1], point[3 * pv + 2]};
    EdgeTerms t;
    double Jx[12], Jp[6];
    linearize_edge(P, P.pm_cam[e], T, p, P.pm_uv[2 * e], P.pm_uv[2 * e + 1], t, Jx, Jp);
    load_edge_weighting(P, P.pm_info, P.pm_delta, e, t);
    double A0[6], A1[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      A0[r] = t.w * (t.o00 * Jx[r] + t.o01 * Jx[6 + r]);
      A1[r] = t.w * (t.o01 * Jx[r] + t.o11 * Jx[6 + r]);
    }
    const double r0 = -t.w * t.we0, r1 = -t.w * t.we1;
#pragma unroll
    for (int r = 0; r < 6; ++r) acc[r] += Jx[r] * r0 + Jx[6 + r] * r1;
    int k = 6;
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
      for (int cc = r; cc < 6; ++cc) acc[k++] += Jx[r] * A0[cc] + Jx[6 + r] * A1[cc];
  }
#pragma unroll
  for (int i = 0; i < 27; ++i) {
    const double v = warp_sum(acc[i]);
    if ((threadIdx.x & 31) == 0) red[i][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x < 27) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kHppThreads / 32; ++w) s += red[threadIdx.x][w];
    P.hpp_part[27 * (size_t)c + threadIdx.x] = s;
  }
}

// fold the chunk partials of every pose in chunk order; keep the diagonal for lambda init
__global__ void k_hpp_reduce(const DeviceProblem P) {
  const Control *ctl = P.ctl;
  if (ctl->done || !ctl->need_linearize) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.n_fp * 27) return;
  const int q = i / 27, k = i % 27;
  double s = 0.0;
  for (int c = P.q_chunk_ptr[q]; c < P.q_chunk_ptr[q + 1]; ++c) s += P.hpp_part[27 * (size_t)c + k];
  P.hpp[i] = s;
  // upper-triangle offsets of the diagonal entries: 6, 12, 17, 21, 24, 26
  const int dsel = k == 6 ? 0 : k == 12 ? 1 : k == 17 ? 2 : k == 21 ? 3 : k == 24 ? 4 : k == 26 ? 5 : -1;
  if (dsel >= 0) P.diag_buf[6 * q + dsel] = s;
}

// max |H_jj| over pose and landmark diagonals (levenberg.cpp:152-166) -> scal[3]
__global__ void __launch_bounds__(256) k_maxdiag(const DeviceProblem P) {
  const Control *ctl = P.ctl;
  if (ctl->done || !ctl->first_iteration) return;
  __shared__ double red[8];
  double m = 0.0;
  for (int i = threadIdx.x; i < 6 * P.n_fp; i += 256) m = fmax(m, fabs(P.diag_buf[i]));
  for (int i = threadIdx.x; i < P.n_lin_blocks; i += 256) m = fmax(m, P.maxdiag_part[i]);
  m = block_max<256>(m, red);
  if (threadIdx.x == 0) P.scal[3] = m;
}

__global__ void k_lambda_init(const DeviceProblem P) {
  Control *ctl = P.ctl;
  if (ctl->done || !ctl->first_iteration) return;
  ctl->maxdiag = P.scal[3];
  ctl->lambda = ctl->user_lambda > 0 ? ctl->user_lambda : ctl->tau * P.scal[3];
  ctl->ni = 2.0;
  ctl->first_iteration = 0;
}

// ---------------------------------------------------------------------------------------------
// k_prepare_system — "Hschur = Hpp" with the lambda of setLambda on the diagonal
// (block_solver.hpp:334-335,524-539), fill blocks zeroed, bschur = b_p (:397).  Written straight
// into the factor storage.  With several ranks each rank contributes its own Hpp part and only
// rank 0 adds lambda; the all-reduce completes the sum.
__global__ void k_prepare_system(const DeviceProblem P) {
  const Control *ctl = P.ctl;
  if (ctl->done) return;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t nL = 36 * (size_t)P.n_blocks;
  if (i < nL) {
    const int b = (int)(i / 36), e = (int)(i % 36);
    const int col = P.blk_col[b];
    double v = 0.0;
    if (P.blk_row[b] == col) {
      const int r = e / 6, c = e % 6;
      const int lo = r < c ? r : c, hi = r < c ? c : r;
      v = P.hpp[27 * (size_t)col + 6 + lo * 6 - lo * (lo - 1) / 2 + (hi - lo)];
      if (r == c && ctl->rank == 0) v += ctl->lambda;
    }
    P.sys[i] = v;
  } else if (i < nL + 6 * (size_t)P.n_fp) {
    const size_t k = i - nL;
    const double v = P.hpp[27 * (k / 6) + (k % 6)];
    P.sys[i] = v;                          // bschur
    P.sys[i + 6 * (size_t)P.n_fp] = v;     // b_p (kept for computeScale)
  }
}

// ---------------------------------------------------------------------------------------------
// k_schur — landmark elimination of BlockSolver::solve (block_solver.hpp:342-400):
// Dinv = (Hll + lambda I)^-1, bschur -= W Dinv b_l, S(i1,i2) -= W_i1 Dinv W_i2^T.
// One thread per free landmark; contributions go to the (L2-resident) reduced system with fp64
// atomic adds.
__global__ void __launch_bounds__(kLinThreads) k_schur(const DeviceProblem P) {
  const Control *ctl = P.ctl;
  if (ctl->done) return;
  const int sl = blockIdx.x * blockDim.x + threadIdx.x;
  if (sl >= P.n_slots || !P.slot_free[sl]) return;
  const double lambda = ctl->lambda;
  double H[6], Di[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) H[i] = P.Hll[6 * (size_t)sl + i];
  H[0] += lambda; H[3] += lambda; H[5] += lambda;
  sym3_inverse(H, Di);
#pragma unroll
  for (int i = 0; i < 6; ++i) P.Dinv[6 * (size_t)sl + i] = Di[i];
  const double b0 = P.bl[3 * (size_t)sl], b1 = P.bl[3 * (size_t)sl + 1], b2 = P.bl[3 * (size_t)sl + 2];
  const double db0 = Di[0] * b0 + Di[1] * b1 + Di[2] * b2;
  const double db1 = Di[1] * b0 + Di[3] * b1 + Di[4] * b2;
  const double db2 = Di[2] * b0 + Di[4] * b1 + Di[5] * b2;
  double *L = P.sys;
  double *bs = P.sys + 36 * (size_t)P.n_blocks;
  const int a0 = P.slot_pair_ptr[sl], a1 = P.slot_pair_ptr[sl + 1];
  int combo = P.slot_combo_ptr[sl];
  for (int a = a0; a < a1; ++a) {
    const int qi = P.pair_q[a];
    if (qi < 0) break;  // fixed-pose pairs come last
    double Wi[18], BD[18];
    const double2 *src = reinterpret_cast<const double2 *>(P.W + 18 * (size_t)a);
#pragma unroll
    for (int i = 0; i < 9; ++i) { const double2 t = src[i]; Wi[2 * i] = t.x; Wi[2 * i + 1] = t.y; }
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      const double w0 = Wi[3 * r], w1 = Wi[3 * r + 1], w2 = Wi[3 * r + 2];
      BD[3 * r + 0] = w0 * Di[0] + w1 * Di[1] + w2 * Di[2];
      BD[3 * r + 1] = w0 * Di[1] + w1 * Di[3] + w2 * Di[4];
      BD[3 * r + 2] = w0 * Di[2] + w1 * Di[4] + w2 * Di[5];
      atomicAdd(bs + 6 * (size_t)qi + r, -(w0 * db0 + w1 * db1 + w2 * db2));
    }
    for (int a2 = a; a2 < a1; ++a2) {
      if (P.pair_q[a2] < 0) break;
      double Wj[18];
      const double2 *s2 = reinterpret_cast<const double2 *>(P.W + 18 * (size_t)a2);
#pragma unroll
      for (int i = 0; i < 9; ++i) { const double2 t = s2[i]; Wj[2 * i] = t.x; Wj[2 * i + 1] = t.y; }
      double *dst = L + 36 * (size_t)P.combo_blk[combo++];  // block (row q_a2, col q_a)
#pragma unroll
      for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int c = 0; c < 6; ++c)
          atomicAdd(dst + 6 * r + c,
                    -(Wj[3 * r] * BD[3 * c] + Wj[3 * r + 1] * BD[3 * c + 1] + Wj[3 * r + 2] * BD[3 * c + 2]));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// k_reduced_solve — the reduced pose system: block-sparse left-looking Cholesky over the
// elimination-tree levels, forward/backward substitution, then the pose part of
// SparseOptimizer::update (sparse_optimizer.cpp:433-446) and of computeScale
// (levenberg.cpp:168-175).  Replaces LinearSolverCSparse::solve
// (g2o/solvers/csparse/linear_solver_csparse.h:106-142) + cs_chol_workspace
// (csparse_extension.cpp:67-122), incl. the "pivot <= 0 => fail" rule (:115).
// Single CTA; a warp per block column inside a level.
__device__ __forceinline__ bool chol6_inplace(double *A) {  // lower, row-major 6x6, one thread
  bool ok = true;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    double d = A[7 * j];
#pragma unroll
    for (int k = 0; k < j; ++k) d -= A[6 * j + k] * A[6 * j + k];
    if (!(d > 0.0)) { ok = false; d = 1.0; }
    const double l = sqrt(d);
    A[7 * j] = l;
    const double il = 1.0 / l;
#pragma unroll
    for (int i = j + 1; i < 6; ++i) {
      double s = A[6 * i + j];
#pragma unroll
      for (int k = 0; k < j; ++k) s -= A[6 * i + k] * A[6 * j + k];
      A[6 * i + j] = s * il;
    }
#pragma unroll
    for (int c = j + 1; c < 6; ++c) A[6 * j + c] = 0.0;
  }
  return ok;
}

__global__ void __launch_bounds__(kSolveThreads) k_reduced_solve(const DeviceProblem P) {
  Control *ctl = P.ctl;
  if (ctl->done) return;
  __shared__ int s_fail;
  __shared__ double red[kSolveThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = kSolveThreads / 32;
  double *L = P.sys;
  double *bs = P.sys + 36 * (size_t)P.n_blocks;
  const double *bp = bs + 6 * (size_t)P.n_fp;
  double *x = P.xp;
  if (tid == 0) s_fail = 0;
  __syncthreads();

  // ---- numeric factorisation, level by level
  for (int lv = 0; lv < P.n_levels; ++lv) {
    for (int t = P.level_ptr[lv] + warp; t < P.level_ptr[lv + 1]; t += NW) {
      const int j = P.level_col[t];
      // 1. gather the updates of column j: L[dst] -= L[a] L[b]^T ; lane owns entries e, e+32
      for (int u = P.upd_ptr[j]; u < P.upd_ptr[j + 1]; ++u) {
        const double *A = L + 36 * (size_t)P.upd_a[u];
        const double *B = L + 36 * (size_t)P.upd_b[u];
        double *D = L + 36 * (size_t)P.upd_dst[u];
        for (int e = lane; e < 36; e += 32) {
          const int r = e / 6, c = e % 6;
          double s = 0.0;
#pragma unroll
          for (int k = 0; k < 6; ++k) s += A[6 * r + k] * B[6 * c + k];
          D[e] -= s;
        }
      }
      __syncwarp();
      // 2. diagonal block
      double *Djj = L + 36 * (size_t)P.col_ptr[j];
      if (lane == 0) {
        double A[36];
#pragma unroll
        for (int i = 0; i < 36; ++i) A[i] = Djj[i];
        if (!chol6_inplace(A)) s_fail = 1;
#pragma unroll
        for (int i = 0; i < 36; ++i) Djj[i] = A[i];
      }
      __syncwarp();
      // 3. sub-diagonal blocks: X L_jj^T = B, one lane per (block, row)
      const int nb = P.col_ptr[j + 1] - P.col_ptr[j] - 1;
      for (int w = lane; w < 6 * nb; w += 32) {
        double *row = L + 36 * (size_t)(P.col_ptr[j] + 1 + w / 6) + 6 * (w % 6);
        double xr[6];
#pragma unroll
        for (int c = 0; c < 6; ++c) {
          double s = row[c];
#pragma unroll
          for (int k = 0; k < c; ++k) s -= xr[k] * Djj[6 * c + k];
          xr[c] = s / Djj[7 * c];
        }
#pragma unroll
        for (int c = 0; c < 6; ++c) row[c] = xr[c];
      }
    }
    __syncthreads();
  }
  const bool fail = s_fail != 0;

  if (!fail) {
    // ---- forward: L y = bschur (rows by level)
    for (int lv = 0; lv < P.n_levels; ++lv) {
      for (int t = P.level_ptr[lv] + warp; t < P.level_ptr[lv + 1]; t += NW) {
        const int j = P.level_col[t];
        // lanes 0..5 own one component each
        double s = lane < 6 ? bs[6 * j + lane] : 0.0;
        if (lane < 6) {
          for (int rr = P.row_ptr[j]; rr < P.row_ptr[j + 1]; ++rr) {
            const double *B = L + 36 * (size_t)P.row_blk[rr] + 6 * lane;
            const double *xk = x + 6 * P.row_col[rr];
#pragma unroll
            for (int c = 0; c < 6; ++c) s -= B[c] * xk[c];
          }
        }
        const double *Djj = L + 36 * (size_t)P.col_ptr[j];
        // 6-step forward substitution across lanes 0..5
        double y = 0.0;
#pragma unroll
        for (int c = 0; c < 6; ++c) {
          const double yc = __shfl_sync(0xffffffffu, s, c) / Djj[7 * c];
          if (lane == c) y = yc;
          if (lane > c && lane < 6) s -= Djj[6 * lane + c] * yc;
        }
        if (lane < 6) x[6 * j + lane] = y;
      }
      __syncthreads();
    }
    // ---- backward: L^T x = y (levels in reverse)
    for (int lv = P.n_levels - 1; lv >= 0; --lv) {
      for (int t = P.level_ptr[lv] + warp; t < P.level_ptr[lv + 1]; t += NW) {
        const int j = P.level_col[t];
        double s = lane < 6 ? x[6 * j + lane] : 0.0;
        if (lane < 6) {
          for (int b = P.col_ptr[j] + 1; b < P.col_ptr[j + 1]; ++b) {
            const double *B = L + 36 * (size_t)b;
            const double *xi = x + 6 * P.blk_row[b];
#pragma unroll
            for (int c = 0; c < 6; ++c) s -= B[6 * c + lane] * xi[c];
          }
        }
        const double *Djj = L + 36 * (size_t)P.col_ptr[j];
        double xv = 0.0;
#pragma unroll
        for (int c = 5; c >= 0; --c) {
          const double xc = __shfl_sync(0xffffffffu, s, c) / Djj[7 * c];
          if (lane == c) xv = xc;
          if (lane < c) s -= Djj[6 * c + lane] * xc;
        }
        if (lane < 6) x[6 * j + lane] = xv;
      }
      __syncthreads();
    }
  } else {
    for (int i = tid; i < 6 * P.n_fp; i += kSolveThreads) x[i] = 0.0;
    __syncthreads();
  }

  // ---- pose part of computeScale and of update(): T <- exp(x) T into the trial buffer
  const int cur = ctl->cur;
  const double lambda = ctl->lambda;
  double sc = 0.0;
  for (int i = tid; i < 6 * P.n_fp; i += kSolveThreads) sc += x[i] * (lambda * x[i] + bp[i]);
  sc = block_sum<kSolveThreads>(sc, red);
  for (int q = tid; q < P.n_fp; q += kSolveThreads) {
    const int kv = P.pose_of_q[q];
    double T[7], d[6], out[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) T[i] = P.pose[cur][7 * kv + i];
#pragma unroll
    for (int i = 0; i < 6; ++i) d[i] = x[6 * q + i];
    pose_oplus(T, d, out);
#pragma unroll
    for (int i = 0; i < 7; ++i) P.pose[cur ^ 1][7 * kv + i] = fail ? T[i] : out[i];
  }
  if (tid == 0) {
    ctl->scale_pose = sc;
    ctl->chol_fail = fail ? 1 : 0;
  }
}

// ---------------------------------------------------------------------------------------------
// k_update — landmark back-substitution x_l = Dinv (b_l - W^T x_p) (block_solver.hpp:420-442),
// p <- p + x_l (g2otypes.hpp:54-59), the landmark part of computeScale, and the trial
// computeActiveErrors + activeRobustChi2 (levenberg.cpp:116-117), one thread per landmark slot.
__global__ void __launch_bounds__(kLinThreads) k_update(const DeviceProblem P) {
  const Control *ctl = P.ctl;
  if (ctl->done) return;
  __shared__ double red[kLinThreads / 32];
  const int cur = ctl->cur;
  const bool fail = ctl->chol_fail != 0;
  const double lambda = ctl->lambda;
  const double *__restrict__ pose_new = P.pose[cur ^ 1];
  const int sl = blockIdx.x * blockDim.x + threadIdx.x;
  double chi = 0.0, scale = 0.0;
  if (sl < P.n_slots) {
    const int pv = P.slot_vertex[sl];
    double p[3] = {P.point[cur][3 * pv], P.point[cur][3 * pv + 1], P.point[cur][3 * pv + 2]};
    const int a0 = P.slot_pair_ptr[sl], a1 = P.slot_pair_ptr[sl + 1];
    if (P.slot_free[sl]) {
      const double b0 = P.bl[3 * (size_t)sl], b1 = P.bl[3 * (size_t)sl + 1], b2 = P.bl[3 * (size_t)sl + 2];
      double c0 = b0, c1 = b1, c2 = b2;
      for (int a = a0; a < a1; ++a) {
        const int q = P.pair_q[a];
        if (q < 0) break;
        const double2 *src = reinterpret_cast<const double2 *>(P.W + 18 * (size_t)a);
        double Wv[18];
#pragma unroll
        for (int i = 0; i < 9; ++i) { const double2 t = src[i]; Wv[2 * i] = t.x; Wv[2 * i + 1] = t.y; }
#pragma unroll
        for (int r = 0; r < 6; ++r) {
          const double xr = P.xp[6 * q + r];
          c0 -= Wv[3 * r] * xr; c1 -= Wv[3 * r + 1] * xr; c2 -= Wv[3 * r + 2] * xr;
        }
      }
      const double *Di = P.Dinv + 6 * (size_t)sl;
      double x0 = Di[0] * c0 + Di[1] * c1 + Di[2] * c2;
      double x1 = Di[1] * c0 + Di[3] * c1 + Di[4] * c2;
      double x2 = Di[2] * c0 + Di[4] * c1 + Di[5] * c2;
      if (fail) { x0 = x1 = x2 = 0.0; }
      scale = x0 * (lambda * x0 + b0) + x1 * (lambda * x1 + b1) + x2 * (lambda * x2 + b2);
      p[0] += x0; p[1] += x1; p[2] += x2;
      P.point[cur ^ 1][3 * pv] = p[0]; P.point[cur ^ 1][3 * pv + 1] = p[1]; P.point[cur ^ 1][3 * pv + 2] = p[2];
    }
    for (int a = a0; a < a1; ++a) {
      const int kv = P.pair_vertex[a];
      double T[7];
#pragma unroll
      for (int i = 0; i < 7; ++i) T[i] = pose_new[7 * kv + i];
      const int e1 = P.pair_edge_ptr[a + 1];
      for (int e = P.pair_edge_ptr[a]; e < e1; ++e) {
        EdgeTerms t;
        edge_error(P.cams.K, P.cams.ext[P.e_cam[e]], T, p, P.e_uv[2 * e], P.e_uv[2 * e + 1], t.e0, t.e1);
        load_edge_weighting(P, P.e_info, P.e_delta, e, t);
        chi += t.rho0;
      }
    }
  }
  const double s = block_sum<kLinThreads>(chi, red);
  const double sc = block_sum<kLinThreads>(scale, red);
  if (threadIdx.x == 0) { P.chi_new_part[blockIdx.x] = s; P.scale_part[blockIdx.x] = sc; }
}

// partial sums -> scal[0..2] = chi(current), chi(trial), landmark part of computeScale
__global__ void __launch_bounds__(256) k_reduce_partials(const DeviceProblem P) {
  const Control *ctl = P.ctl;
  if (ctl->done) return;
  __shared__ double red[8];
  double a = 0.0, b = 0.0, c = 0.0;
  for (int i = threadIdx.x; i < P.n_lin_blocks; i += 256) a += P.chi_cur_part[i];
  for (int i = threadIdx.x; i < P.n_upd_blocks; i += 256) { b += P.chi_new_part[i]; c += P.scale_part[i]; }
  a = block_sum<256>(a, red);
  b = block_sum<256>(b, red);
  c = block_sum<256>(c, red);
  if (threadIdx.x == 0) { P.scal[0] = a; P.scal[1] = b; P.scal[2] = c; }
}

// ---------------------------------------------------------------------------------------------
// k_control — the accept/reject law of OptimizationAlgorithmLevenberg::solve
// (levenberg.cpp:119-149) and the iteration bookkeeping of SparseOptimizer::optimize
// (sparse_optimizer.cpp:386-426), on the device so that no host round trip sits between trials.
__global__ void k_control(const DeviceProblem P) {
  Control *c = P.ctl;
  if (c->done) return;
  const double currentChi = P.scal[0];
  double tempChi = P.scal[1];
  if (c->outer_iter == 0 && c->qmax == 0) c->chi2_initial = currentChi;
  const bool fail = c->chol_fail != 0;
  if (fail) { tempChi = DBL_MAX; c->cholesky_failures++; }
  double scale = P.scal[2] + c->scale_pose;
  scale += 1e-3;
  double rho = (currentChi - tempChi) / scale;
  if (fail) rho = -1.0;  // a failed factorisation always rejects the step (levenberg.cpp:120-121)
  bool broke = false, accepted = false;
  if (rho > 0 && isfinite(tempChi)) {
    double alpha = 1. - pow((2 * rho - 1), 3);
    alpha = fmin(alpha, c->good_upper);
    const double scaleFactor = fmax(c->good_lower, alpha);
    c->lambda *= scaleFactor;
    c->ni = 2;
    c->cur ^= 1;  // discardTop(): the trial buffer becomes the estimate
    accepted = true;
  } else {
    c->lambda *= c->ni;
    c->ni *= 2;  // pop(): the estimate buffer is simply kept
    if (!isfinite(c->lambda)) broke = true;
  }
  if (!broke) c->qmax++;
  c->rho = rho;
  c->temp_chi = tempChi;
  c->current_chi = accepted ? tempChi : currentChi;
  const bool again = !broke && rho < 0 && c->qmax < c->max_trials;
  if (again) {
    c->need_linearize = 0;
    return;
  }
  const int result = (c->qmax == c->max_trials || rho == 0 || !isfinite(c->lambda))
                         ? SSBA_SOLVER_TERMINATE : SSBA_SOLVER_OK;
  if (c->n_records < SSBA_MAX_ITER_RECORDS) {
    ssba_iter_record &r = c->records[c->n_records++];
    r.chi2 = c->current_chi; r.lambda = c->lambda; r.trials = c->qmax; r.result = result;
  }
  c->last_result = result;
  c->outer_iter++;
  c->qmax = 0;
  c->need_linearize = 1;
  if (result != SSBA_SOLVER_OK || c->outer_iter >= c->max_iters) c->done = 1;
}

// ---------------------------------------------------------------------------------------------
// final read-outs at the current estimate: activeChi2 / activeRobustChi2
// (sparse_optimizer.cpp:92-116), the outlier count of backend.cpp:180-197, per-edge errors.
__global__ void __launch_bounds__(kLinThreads) k_final_chi2(const DeviceProblem P, double threshold,
                                                            int write_errors) {
  __shared__ double red[kLinThreads / 32];
  const int cur = P.ctl->cur;
  const double *__restrict__ pose = P.pose[cur];
  const double *__restrict__ point = P.point[cur];
  const int sl = blockIdx.x * blockDim.x + threadIdx.x;
  double plain = 0.0, robust = 0.0, nout = 0.0, nin = 0.0;
  if (sl < P.n_slots) {
    const int pv = P.slot_vertex[sl];
    const double p[3] = {point[3 * pv], point[3 * pv + 1], point[3 * pv + 2]};
    for (int a = P.slot_pair_ptr[sl]; a < P.slot_pair_ptr[sl + 1]; ++a) {
      const int kv = P.pair_vertex[a];
      double T[7];
#pragma unroll
      for (int i = 0; i < 7; ++i) T[i] = pose[7 * kv + i];
      for (int e = P.pair_edge_ptr[a]; e < P.pair_edge_ptr[a + 1]; ++e) {
        EdgeTerms t;
        edge_error(P.cams.K, P.cams.ext[P.e_cam[e]], T, p, P.e_uv[2 * e], P.e_uv[2 * e + 1], t.e0, t.e1);
        load_edge_weighting(P, P.e_info, P.e_delta, e, t);
        plain += t.chi; robust += t.rho0;
        if (t.chi > threshold) nout += 1.0; else nin += 1.0;
        if (write_errors) {
          const int o = P.e_orig[e];
          P.err_out[2 * (size_t)o] = t.e0; P.err_out[2 * (size_t)o + 1] = t.e1;
        }
      }
    }
  }
  const double a = block_sum<kLinThreads>(plain, red);
  const double b = block_sum<kLinThreads>(robust, red);
  const double c = block_sum<kLinThreads>(nout, red);
  const double d = block_sum<kLinThreads>(nin, red);
  if (threadIdx.x == 0) {
    // the partial arrays of the LM loop are free at this point
    P.chi_cur_part[blockIdx.x] = a; P.maxdiag_part[blockIdx.x] = b;
    P.chi_new_part[blockIdx.x] = c; P.scale_part[blockIdx.x] = d;
  }
}

__global__ void __launch_bounds__(256) k_final_reduce(const DeviceProblem P) {
  __shared__ double red[8];
  double a = 0.0, b = 0.0, c = 0.0, d = 0.0;
  for (int i = threadIdx.x; i < P.n_lin_blocks; i += 256) {
    a += P.chi_cur_part[i]; b += P.maxdiag_part[i]; c += P.chi_new_part[i]; d += P.scale_part[i];
  }
  a = block_sum<256>(a, red); b = block_sum<256>(b, red);
  c = block_sum<256>(c, red); d = block_sum<256>(d, red);
  if (threadIdx.x == 0) { P.chi_out[0] = a; P.chi_out[1] = b; P.chi_out[2] = c; P.chi_out[3] = d; }
}

inline int div_up(long long a, int b) { return (int)((a + b - 1) / b); }

}  // namespace

// ------------------------------------------------------------------------------------------------

int kernels_per_linearize() { return 3; }

void launch_linearize(const DeviceProblem &P, cudaStream_t st) {
  k_linearize<<<P.n_lin_blocks, kLinThreads, 0, st>>>(P);
  if (P.n_chunks > 0) k_hpp<<<P.n_chunks, kHppThreads, 0, st>>>(P);
  if (P.n_fp > 0) k_hpp_reduce<<<div_up(27LL * P.n_fp, 128), 128, 0, st>>>(P);
}

void launch_maxdiag(const DeviceProblem &P, cudaStream_t st) { k_maxdiag<<<1, 256, 0, st>>>(P); }

void launch_lambda_init(const DeviceProblem &P, cudaStream_t st) { k_lambda_init<<<1, 1, 0, st>>>(P); }

void launch_prepare_system(const DeviceProblem &P, cudaStream_t st) {
  const long long n = 36LL * P.n_blocks + 6LL * P.n_fp;
  if (n > 0) k_prepare_system<<<div_up(n, 256), 256, 0, st>>>(P);
}

void launch_schur(const DeviceProblem &P, cudaStream_t st) {
  k_schur<<<P.n_lin_blocks, kLinThreads, 0, st>>>(P);
}

void launch_reduced_solve(const DeviceProblem &P, cudaStream_t st) {
  k_reduced_solve<<<1, kSolveThreads, 0, st>>>(P);
}

void launch_update(const DeviceProblem &P, cudaStream_t st) {
  k_update<<<P.n_upd_blocks, kLinThreads, 0, st>>>(P);
}

void launch_reduce_partials(const DeviceProblem &P, cudaStream_t st) {
  k_reduce_partials<<<1, 256, 0, st>>>(P);
}

void launch_control(const DeviceProblem &P, cudaStream_t st) { k_control<<<1, 1, 0, st>>>(P); }

void launch_final_chi2(const DeviceProblem &P, double threshold, cudaStream_t st) {
  k_final_chi2<<<P.n_lin_blocks, kLinThreads, 0, st>>>(P, threshold, 0);
  k_final_reduce<<<1, 256, 0, st>>>(P);
}

void launch_edge_errors(const DeviceProblem &P, cudaStream_t st) {
  k_final_chi2<<<P.n_lin_blocks, kLinThreads, 0, st>>>(P, 0.0, 1);
}

}  // namespace ssba
