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
#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cstdlib>
#include <mutex>

#include "ssba_device.hpp"
#include "ssba_solver_layout.hpp"

namespace ssba {

namespace {

constexpr int kLinThreads = 128;
// = kLinPairs of the host (ssba_structure.cpp): the two must be changed together.  (64 pairs per chunk, tried in
// round 2 for a finer tail of the two CTA waves: update 32.4 -> 33.0 us, schur + fold 28.5 -> 33.9 us - not kept.)
static_assert(kLinThreads == 128, "kLinThreads must equal kLinPairs of ssba_structure.cpp");
#ifndef SSBA_LIN_MINB
#define SSBA_LIN_MINB 3  // CTAs per SM k_linearize is compiled for (factored path: 134 registers, no spills)
#endif

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

// ---- TMA staging of the pose array (cp.async.bulk + mbarrier): the poses of a window are a few KB that every
// pair of a CTA gathers from; one bulk copy per CTA puts them into shared memory while the CTA does other work.
__device__ __forceinline__ unsigned ek_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ek_stage_begin(unsigned long long *bar, double *dst, const double *src, unsigned bytes) {
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ek_smem_u32(bar)), "r"(1u) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ek_smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(ek_smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(ek_smem_u32(bar)) : "memory");
  }
}
// every thread that is going to read the staged data waits (after a __syncthreads that follows ek_stage_begin)
__device__ __forceinline__ void ek_stage_wait(unsigned long long *bar) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(ek_smem_u32(bar)), "r"(0u) : "memory");
}
// bytes of the staged pose array, 0 = too large for shared memory beside the linearisation scratch (then the pairs
// gather from global memory as before)
constexpr unsigned kPoseStageMaxBytes = 32 * 1024;
}  // namespace
unsigned pose_stage_bytes(int n_poses) {
  const unsigned b = (56u * (unsigned)n_poses + 15u) & ~15u;  // the arena pads every array: reading up to 8 bytes on is safe
  return b <= kPoseStageMaxBytes ? b : 0u;
}
namespace {

__device__ __forceinline__ long long gtime_early() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

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

template <int kMode>
__device__ __forceinline__ void linearize_edge(const DeviceProblem &P, int cam, const double *T,
                                               const double *p, double u, double v, EdgeTerms &t,
                                               double *Jx, double *Jp) {
  if (kMode == SSBA_JACOBIAN_NUMERIC)
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
//
// One thread per (pose, landmark) PAIR (its 1..n_cams edges), one CTA per chunk of whole
// landmarks with <= 128 pairs: W is produced in registers and stored as consecutive 144-byte
// records by consecutive threads; the Hll / b_l partials of a landmark's pairs meet in shared
// memory and are folded in pair order by the landmark's first thread.  No atomics, deterministic.
struct PairLin {
  double W[18], H[6], b[3], chi;
};
__device__ __forceinline__ PairRec load_pair_rec(const DeviceProblem &P, int a) {
  const int4 *src = reinterpret_cast<const int4 *>(P.pair_rec + a);
  const int4 x = src[0], y = src[1];
  PairRec r;
  r.pose_row = x.x; r.q = x.y; r.slot = x.z; r.e0 = x.w; r.n_edges = y.x; r.point_row = y.y; r.lfree = y.z; r.pad = 0;
  return r;
}

// `hp` (27 doubles, the caller's own row of shared memory or a global partial) receives the pose
// side: b_p[6], then the upper triangle of J_xi^T (rho' Omega) J_xi [21]
template <int kMode>
__device__ __forceinline__ void linearize_pair(const DeviceProblem &P, double *Wout, int a, const PairRec &rec, const double *pose,
                                               const double *p, PairLin &o, double *hp) {
  const int kv = rec.pose_row;
  const bool lfree = rec.lfree != 0, pfree = rec.q >= 0;
  const bool wpair = lfree && pfree;
  double T[7];
#pragma unroll
  for (int i = 0; i < 7; ++i) T[i] = pose[7 * kv + i];
#pragma unroll
  for (int i = 0; i < 27; ++i) hp[i] = 0.0;
#pragma unroll
  for (int i = 0; i < 18; ++i) o.W[i] = 0.0;
#pragma unroll
  for (int i = 0; i < 6; ++i) o.H[i] = 0.0;
  o.b[0] = o.b[1] = o.b[2] = 0.0;
  o.chi = 0.0;
  const int e1 = rec.e0 + rec.n_edges;
  for (int e = rec.e0; e < e1; ++e) {
    EdgeTerms t;
    double Jx[12], Jp[6];
    linearize_edge<kMode>(P, P.e_cam[e], T, p, P.e_uv[2 * e], P.e_uv[2 * e + 1], t, Jx, Jp);
    load_edge_weighting(P, P.e_info, P.e_delta, e, t);
    o.chi += t.rho0;
    if (pfree) {
      // pose side (base_binary_edge.hpp:104-110): b_i += J_xi^T (-rho' Omega e), Hpp_ii += J_xi^T (rho' Omega) J_xi
      double A0[6], A1[6];
#pragma unroll
      for (int r = 0; r < 6; ++r) {
        A0[r] = t.w * (t.o00 * Jx[r] + t.o01 * Jx[6 + r]);
        A1[r] = t.w * (t.o01 * Jx[r] + t.o11 * Jx[6 + r]);
      }
      const double r0 = -t.w * t.we0, r1 = -t.w * t.we1;
#pragma unroll
      for (int r = 0; r < 6; ++r) hp[r] += Jx[r] * r0 + Jx[6 + r] * r1;
      int k = 6;
#pragma unroll
      for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int cc = r; cc < 6; ++cc) hp[k++] += Jx[r] * A0[cc] + Jx[6 + r] * A1[cc];
    }
    if (lfree) {
      // rows of (rho' Omega) J_p
      double A0[3], A1[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        A0[c] = t.w * (t.o00 * Jp[c] + t.o01 * Jp[3 + c]);
        A1[c] = t.w * (t.o01 * Jp[c] + t.o11 * Jp[3 + c]);
      }
      o.H[0] += Jp[0] * A0[0] + Jp[3] * A1[0];
      o.H[1] += Jp[0] * A0[1] + Jp[3] * A1[1];
      o.H[2] += Jp[0] * A0[2] + Jp[3] * A1[2];
      o.H[3] += Jp[1] * A0[1] + Jp[4] * A1[1];
      o.H[4] += Jp[1] * A0[2] + Jp[4] * A1[2];
      o.H[5] += Jp[2] * A0[2] + Jp[5] * A1[2];
      const double r0 = -t.w * t.we0, r1 = -t.w * t.we1;  // omega_r * rho'
#pragma unroll
      for (int c = 0; c < 3; ++c) o.b[c] += Jp[c] * r0 + Jp[3 + c] * r1;
      if (wpair) {
#pragma unroll
        for (int r = 0; r < 6; ++r)
#pragma unroll
          for (int c = 0; c < 3; ++c) o.W[3 * r + c] += Jx[r] * A0[c] + Jx[6 + r] * A1[c];
      }
    }
  }
  if (wpair) {
    double2 *dst = reinterpret_cast<double2 *>(Wout + 18 * (size_t)a);
#pragma unroll
    for (int i = 0; i < 9; ++i) dst[i] = make_double2(o.W[2 * i], o.W[2 * i + 1]);
  }
}

// Closed-form path (the default).  The quadratic form of a (pose, landmark) pair factors through
// the body-frame point b = T p: for every camera J_xi = Jc [ I | -[b]x ] and J_p = Jc R with
// Jc = Jpi R_ext (2x3), so with  M = sum_cam Jc^T (rho' Omega) Jc  (3x3) and
// g = sum_cam Jc^T (-rho' Omega e)  the pair contributes
//     Hpp = A^T M A,  b_p = A^T g,  Hll = R^T M R,  b_l = R^T g,  W = A^T M R,   A = [ I | -[b]x ],
// i.e. the per-edge work is a 3x3 accumulation and the 6x6 / 6x3 / 3x3 blocks are expanded once per
// pair (about a quarter of the fp64 work of expanding every edge, and nothing is re-read from shared
// memory).  Same sums as base_binary_edge.hpp:61-134 up to the order of the additions.
__device__ __forceinline__ void cross3(const double *a, double x, double y, double z, double &ox, double &oy, double &oz) {
  ox = a[1] * z - a[2] * y; oy = a[2] * x - a[0] * z; oz = a[0] * y - a[1] * x;
}
__device__ __forceinline__ void linearize_pair_factored(const DeviceProblem &P, double *Wout, int a, const PairRec &rec, const double *pose,
                                                        const double *p, PairLin &o, double *hp) {
  const int kv = rec.pose_row;
  const bool lfree = rec.lfree != 0, pfree = rec.q >= 0;
  const bool wpair = lfree && pfree;
  double T[7];
#pragma unroll
  for (int i = 0; i < 7; ++i) T[i] = pose[7 * kv + i];
  double b[3];
  se3_act(T, p[0], p[1], p[2], b[0], b[1], b[2]);
  double M[6] = {0, 0, 0, 0, 0, 0}, g[3] = {0, 0, 0};  // M: xx xy xz yy yz zz
  o.chi = 0.0;
  const int e1 = rec.e0 + rec.n_edges;
  for (int e = rec.e0; e < e1; ++e) {
    const int cam = P.e_cam[e];
    const double *K = P.cams.K, *Re = P.ext_R[cam];
    double cx, cy, cz;
    se3_act(P.cams.ext[cam], b[0], b[1], b[2], cx, cy, cz);
    const double n0 = K[0] * cx + K[1] * cy + K[2] * cz;
    const double n1 = K[3] * cx + K[4] * cy + K[5] * cz;
    const double dn = K[6] * cx + K[7] * cy + K[8] * cz;
    const double id = 1.0 / dn, id2 = id * id;
    EdgeTerms t;
    t.e0 = P.e_uv[2 * e] - n0 / dn;
    t.e1 = P.e_uv[2 * e + 1] - n1 / dn;
    load_edge_weighting(P, P.e_info, P.e_delta, e, t);
    o.chi += t.rho0;
    double J0[3], J1[3];  // rows of Jc = Jpi R_ext
    {
      double q0[3], q1[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        q0[c] = -(K[c] * id - n0 * K[6 + c] * id2);
        q1[c] = -(K[3 + c] * id - n1 * K[6 + c] * id2);
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        J0[c] = q0[0] * Re[c] + q0[1] * Re[3 + c] + q0[2] * Re[6 + c];
        J1[c] = q1[0] * Re[c] + q1[1] * Re[3 + c] + q1[2] * Re[6 + c];
      }
    }
    double A0[3], A1[3];  // rows of (rho' Omega) Jc
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      A0[c] = t.w * (t.o00 * J0[c] + t.o01 * J1[c]);
      A1[c] = t.w * (t.o01 * J0[c] + t.o11 * J1[c]);
    }
    M[0] += J0[0] * A0[0] + J1[0] * A1[0]; M[1] += J0[0] * A0[1] + J1[0] * A1[1]; M[2] += J0[0] * A0[2] + J1[0] * A1[2];
    M[3] += J0[1] * A0[1] + J1[1] * A1[1]; M[4] += J0[1] * A0[2] + J1[1] * A1[2]; M[5] += J0[2] * A0[2] + J1[2] * A1[2];
    const double r0 = -t.w * t.we0, r1 = -t.w * t.we1;
#pragma unroll
    for (int c = 0; c < 3; ++c) g[c] += J0[c] * r0 + J1[c] * r1;
  }
  const double Mr[3][3] = {{M[0], M[1], M[2]}, {M[1], M[3], M[4]}, {M[2], M[4], M[5]}};
  // pose side: Hpp = [[M, N], [N^T, b x N]], N(r, :) = b x M(r, :);  b_p = [g ; b x g]
  if (pfree) {
    double N[3][3], BR[3][3];
#pragma unroll
    for (int r = 0; r < 3; ++r) cross3(b, Mr[r][0], Mr[r][1], Mr[r][2], N[r][0], N[r][1], N[r][2]);
#pragma unroll
    for (int c = 0; c < 3; ++c) cross3(b, N[0][c], N[1][c], N[2][c], BR[0][c], BR[1][c], BR[2][c]);
    hp[0] = g[0]; hp[1] = g[1]; hp[2] = g[2];
    cross3(b, g[0], g[1], g[2], hp[3], hp[4], hp[5]);
    int k = 6;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
      for (int c = r; c < 3; ++c) hp[k++] = Mr[r][c];
#pragma unroll
      for (int c = 0; c < 3; ++c) hp[k++] = N[r][c];
    }
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = r; c < 3; ++c) hp[k++] = BR[r][c];
  } else {
#pragma unroll
    for (int i = 0; i < 27; ++i) hp[i] = 0.0;
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) o.H[i] = 0.0;
  o.b[0] = o.b[1] = o.b[2] = 0.0;
  if (lfree) {
    double R[9], MR[3][3];
    quat_to_matrix(T, R);
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) MR[r][c] = Mr[r][0] * R[c] + Mr[r][1] * R[3 + c] + Mr[r][2] * R[6 + c];
    // Hll = R^T (M R), b_l = R^T g
    o.H[0] = R[0] * MR[0][0] + R[3] * MR[1][0] + R[6] * MR[2][0];
    o.H[1] = R[0] * MR[0][1] + R[3] * MR[1][1] + R[6] * MR[2][1];
    o.H[2] = R[0] * MR[0][2] + R[3] * MR[1][2] + R[6] * MR[2][2];
    o.H[3] = R[1] * MR[0][1] + R[4] * MR[1][1] + R[7] * MR[2][1];
    o.H[4] = R[1] * MR[0][2] + R[4] * MR[1][2] + R[7] * MR[2][2];
    o.H[5] = R[2] * MR[0][2] + R[5] * MR[1][2] + R[8] * MR[2][2];
#pragma unroll
    for (int c = 0; c < 3; ++c) o.b[c] = R[c] * g[0] + R[3 + c] * g[1] + R[6 + c] * g[2];
    if (wpair) {
      // W = [M R ; b x (M R)]  (6x3, row-major)
      double Wv[18];
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) Wv[3 * r + c] = MR[r][c];
#pragma unroll
      for (int c = 0; c < 3; ++c) cross3(b, MR[0][c], MR[1][c], MR[2][c], Wv[9 + c], Wv[12 + c], Wv[15 + c]);
      double2 *dst = reinterpret_cast<double2 *>(Wout + 18 * (size_t)a);
#pragma unroll
      for (int i = 0; i < 9; ++i) dst[i] = make_double2(Wv[2 * i], Wv[2 * i + 1]);
    }
  }
}

// shared memory of one chunk's linearisation (k_linearize, and k_update when it linearises the trial state)
struct LinSmem {
  double part[kLinThreads][9];
  double hp[kLinThreads][27];
  double red[kLinThreads / 32];
  int lp_ptr[kLinThreads + 1];
  uint8_t lp_pair[kLinThreads];
};

// The linearisation of chunk `blockIdx.x` at the state (pose, point) into the buffers `lin`: W, Hll, b_l, the
// Hpp / b_p partials of the chunk's poses; returns this thread's share of the robust chi2 and of max |Hll_jj|.
template <int kMode>
// (`point` is NOT restrict / read-only: k_update reads positions its own CTA has just written)
__device__ __forceinline__ void linearize_chunk(const DeviceProblem &P, int lin, const double *__restrict__ pose,
                                                const double *point, LinSmem &sm, double &chi, double &mx) {
  double *Wout = P.W[lin], *Hll = P.Hll[lin], *bl = P.bl[lin], *hpp_part = P.hpp_part[lin];
  const int s0 = P.lchunk_slot[blockIdx.x], s1 = P.lchunk_slot[blockIdx.x + 1];
  const int lp0 = P.lchunk_lp_ptr[blockIdx.x], lp1 = P.lchunk_lp_ptr[blockIdx.x + 1];
  const int a0 = P.slot_pair_ptr[s0], a1 = P.slot_pair_ptr[s1];
  const int tid = threadIdx.x;
  chi = 0.0; mx = 0.0;
  if (a1 - a0 <= kLinThreads) {
    // the chunk's local-pose lists, fetched now and used after the barrier (their latency hides
    // behind the linearisation)
    const int lpp0 = P.lp_pair_ptr[lp0];
    if (tid <= lp1 - lp0) sm.lp_ptr[tid] = P.lp_pair_ptr[lp0 + tid] - lpp0;
    if (tid < P.lp_pair_ptr[lp1] - lpp0) sm.lp_pair[tid] = P.lp_pair[lpp0 + tid];
    if (tid < a1 - a0) {
      const int a = a0 + tid;
      const PairRec rec = load_pair_rec(P, a);
      const int pv = rec.point_row;
      const double p[3] = {point[3 * pv], point[3 * pv + 1], point[3 * pv + 2]};
      PairLin o;
      if (kMode == SSBA_JACOBIAN_ANALYTIC) linearize_pair_factored(P, Wout, a, rec, pose, p, o, sm.hp[tid]);
      else linearize_pair<kMode>(P, Wout, a, rec, pose, p, o, sm.hp[tid]);
      chi = o.chi;
#pragma unroll
      for (int i = 0; i < 6; ++i) sm.part[tid][i] = o.H[i];
#pragma unroll
      for (int i = 0; i < 3; ++i) sm.part[tid][6 + i] = o.b[i];
    }
    __syncthreads();
    // pose side: fold the pairs of every distinct pose of this chunk, in pair order
    for (int w = tid; w < 27 * (lp1 - lp0); w += kLinThreads) {
      const int lp = lp0 + w / 27, k = w % 27;
      double acc = 0.0;
      for (int i = sm.lp_ptr[lp - lp0]; i < sm.lp_ptr[lp - lp0 + 1]; ++i) acc += sm.hp[sm.lp_pair[i]][k];
      hpp_part[27 * (size_t)lp + k] = acc;
    }
    const int sl = s0 + tid;
    if (sl < s1 && P.slot_free[sl]) {
      double acc[9];
#pragma unroll
      for (int i = 0; i < 9; ++i) acc[i] = 0.0;
      for (int a = P.slot_pair_ptr[sl] - a0; a < P.slot_pair_ptr[sl + 1] - a0; ++a) {
#pragma unroll
        for (int i = 0; i < 9; ++i) acc[i] += sm.part[a][i];
      }
#pragma unroll
      for (int i = 0; i < 6; ++i) Hll[6 * (size_t)sl + i] = acc[i];
#pragma unroll
      for (int i = 0; i < 3; ++i) bl[3 * (size_t)sl + i] = acc[6 + i];
      mx = fmax(fabs(acc[0]), fmax(fabs(acc[3]), fabs(acc[5])));
    }
  } else {
    // a single landmark seen by more than 128 poses: threads stride over its pairs
    const int sl = s0;
    const int pv = P.slot_vertex[sl];
    const bool lfree = P.slot_free[sl] != 0;
    const double p[3] = {point[3 * pv], point[3 * pv + 1], point[3 * pv + 2]};
    double acc[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) acc[i] = 0.0;
    // every pair of one landmark is on a different pose: one Hpp partial per free-pose pair, in
    // pair order (the partial index advances with the free-pose pairs before this one)
    for (int a = a0 + tid; a < a1; a += kLinThreads) {
      PairLin o;
      // free-pose pairs come first inside a landmark, so the rank of one is simply a - a0
      const PairRec rec = load_pair_rec(P, a);
      double *hp_dst = rec.q >= 0 ? hpp_part + 27 * (size_t)(lp0 + (a - a0)) : sm.hp[tid];
      if (kMode == SSBA_JACOBIAN_ANALYTIC) linearize_pair_factored(P, Wout, a, rec, pose, p, o, hp_dst);
      else linearize_pair<kMode>(P, Wout, a, rec, pose, p, o, hp_dst);
      chi += o.chi;
#pragma unroll
      for (int i = 0; i < 6; ++i) acc[i] += o.H[i];
#pragma unroll
      for (int i = 0; i < 3; ++i) acc[6 + i] += o.b[i];
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) acc[i] = block_sum<kLinThreads>(acc[i], sm.red);
    if (tid == 0 && lfree) {
#pragma unroll
      for (int i = 0; i < 6; ++i) Hll[6 * (size_t)sl + i] = acc[i];
#pragma unroll
      for (int i = 0; i < 3; ++i) bl[3 * (size_t)sl + i] = acc[6 + i];
      mx = fmax(fabs(acc[0]), fmax(fabs(acc[3]), fabs(acc[5])));
    }
  }
}

template <int kMode>
// measured on B200 (cfg3): per-edge expansion 29 us (168 registers, spills); factored per pair 23 us at three
// CTAs per SM, 24 us at four (125 registers) - the kernel is bound by its dependent-load chains, not occupancy
__global__ void __launch_bounds__(kLinThreads, SSBA_LIN_MINB) k_linearize(const DeviceProblem P) {
  const Control *ctl = P.ctl;
  if (ctl->done || !ctl->need_linearize) return;
  __shared__ LinSmem sm;
  __shared__ __align__(8) unsigned long long s_bar;
  extern __shared__ __align__(16) double s_pose[];
  const int cur = ctl->cur;
  const unsigned pb = P.pose_stage;
  const double *pose = P.pose[cur];
  if (pb) {
    ek_stage_begin(&s_bar, s_pose, pose, pb);
    __syncthreads();
    ek_stage_wait(&s_bar);
    pose = s_pose;
  }
  double chi, mx;
  linearize_chunk<kMode>(P, ctl->lin, pose, P.point[cur], sm, chi, mx);
  const double s_ = block_sum<kLinThreads>(chi, sm.red);
  const double m = block_max<kLinThreads>(mx, sm.red);
  if (threadIdx.x == 0) { P.chi_cur_part[blockIdx.x] = s_; P.maxdiag_part[blockIdx.x] = m; }
}

// Hpp / b_p of pose q (entry k: 0..5 = b, 6..26 = upper triangle): the partials the linearize CTAs
// wrote for the pose, folded in chunk order (deterministic).  One warp per pose, lane k < 27 owns
// entry k; the loads of sixteen partials are in flight together, the adds stay in list order.
__device__ __forceinline__ double warp_fold_pose(const DeviceProblem &P, const double *__restrict__ hpp_part, int q, int lane) {
  double s = 0.0;
  const int i0 = P.q_part_ptr[q], i1 = P.q_part_ptr[q + 1];
  for (int base = i0; base < i1; base += 32) {
    const int cnt = i1 - base < 32 ? i1 - base : 32;
    const int mine = lane < cnt ? P.q_part[base + lane] : 0;  // 32 list entries with one coalesced load
    for (int j0 = 0; j0 < cnt; j0 += 16) {
      double v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int part = __shfl_sync(0xffffffffu, mine, (j0 + j) & 31);
        v[j] = (j0 + j < cnt && lane < 27) ? hpp_part[27 * (size_t)part + lane] : 0.0;
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) s += v[j];  // list order; the padding adds exact zeros
    }
  }
  return s;
}
// upper-triangle offset of diagonal entry d of the 6x6 block: 6, 12, 17, 21, 24, 26
__device__ __forceinline__ int hpp_diag_index(int d) { return 6 + d * 6 - d * (d - 1) / 2; }

// first slot of an optimize()/step() call: every pose's Hpp partials folded once -> hpp_fold, and
// the Hpp diagonals -> diag_buf for computeLambdaInit (summed over ranks by the host-enqueued
// all-reduce before k_maxdiag when there are several).  One warp per pose.
__global__ void __launch_bounds__(128) k_fold(const DeviceProblem P) {
  const Control *ctl = P.ctl;
  if (ctl->done || !(ctl->need_linearize || ctl->need_fold)) return;
  const int q = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (q >= P.n_fp) return;
  const double s = warp_fold_pose(P, P.hpp_part[ctl->lin], q, lane);
  if (lane < 27) P.hpp_fold[27 * q + lane] = s;
#pragma unroll
  for (int d = 0; d < 6; ++d) if (lane == hpp_diag_index(d)) P.diag_buf[6 * q + d] = s;
}

// max |H_jj| over pose and landmark diagonals (levenberg.cpp:152-166) -> scal[3]
__global__ void __launch_bounds__(1024) k_maxdiag(const DeviceProblem P) {
  const Control *ctl = P.ctl;
  if (ctl->done || !ctl->first_iteration) return;
  __shared__ double red[32];
  double m = 0.0;
  for (int i = threadIdx.x; i < 6 * P.n_fp; i += 1024) m = fmax(m, fabs(P.diag_buf[i]));
  for (int i = threadIdx.x; i < P.n_lin_blocks; i += 1024) m = fmax(m, P.maxdiag_part[i]);
  m = block_max<1024>(m, red);
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
// k_schur — the reduced system of BlockSolver::solve (block_solver.hpp:334-400), accumulated with
// fp64 atomic adds into the factor storage `sys`, which the previous trial's k_update left zeroed:
//   "Hschur = Hpp" with the lambda of setLambda on the diagonal (:334-335,524-539), bschur = b_p (:397)
//       — the first CTAs, one warp per pose: fold the pose's Hpp partials (once per linearisation;
//       a re-trial with a new lambda reads the folded copy) and add the diagonal block, bschur,
//       and store b_p.  With several ranks each rank contributes its own Hpp part and only rank
//       0 adds lambda; the all-reduce completes the sum.
//   landmark elimination (:342-400): Dinv = (Hll + lambda I)^-1, bschur -= W Dinv b_l,
//       S(i1,i2) -= W_i1 Dinv W_i2^T — the other CTAs, one warp per work unit = a run of <= 16
//       landmarks seen by the same k free poses (the host sorts the landmarks so that such runs
//       exist) x <= 32 of the k(k+1)/2 block pairs.  Lane i first inverts the 3x3 block of
//       landmark i of the run into shared memory; then a lane owns ONE 6x6 block pair (a, b) and
//       walks every G-th landmark of the run (G = 32 / block pairs lane groups share the run),
//       accumulating its block in 36 registers; the G partial blocks meet through shuffles and
//       only the per-unit totals go to the L2-resident reduced system.
constexpr int kSchurWarps = 4;
constexpr int kSchurRunPairs = 80;  // W blocks of one run staged in shared memory (= host kSchurRunPairs)
constexpr int kSchurRun = 16;       // landmarks per run (= host kSchurRun)

// where k_schur accumulates: the solver's system, or (peer-memory exchange) this rank's partial of the
// running trial, which k_exchange_sys then sums over the ranks into the solver's system
__device__ __forceinline__ double *schur_target(const DeviceProblem &P, const Control *ctl) {
  if (!P.use_p2p) return P.sys;
  return reinterpret_cast<double *>(P.peer[ctl->rank] + sizeof(PeerHeader)) + (size_t)((ctl->trial_seq + 1) & 1) * P.sys_doubles;
}

__device__ __forceinline__ void schur_pose_warp(const DeviceProblem &P, const Control *ctl, int q, int lane,
                                                bool prefolded) {
  double s;
  if ((ctl->need_linearize || ctl->need_fold) && !prefolded) {
    s = warp_fold_pose(P, P.hpp_part[ctl->lin], q, lane);
    if (lane < 27) P.hpp_fold[27 * q + lane] = s;
  } else {
    s = lane < 27 ? P.hpp_fold[27 * q + lane] : 0.0;
  }
  if (P.deterministic) return;  // k_schur_reduce puts Hpp + lambda I, bschur and b_p in place
  const double lambda = ctl->rank == 0 ? ctl->lambda : 0.0;
  double *sysacc = schur_target(P, ctl);
  double *D = sysacc + 36 * (size_t)P.col_diag[q];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int e = lane + 32 * h, ee = e < 36 ? e : 0;
    const int r = ee / 6, c = ee - 6 * r;
    const int lo = r < c ? r : c, hi = r < c ? c : r;
    const double v = __shfl_sync(0xffffffffu, s, 6 + lo * 6 - lo * (lo - 1) / 2 + (hi - lo));
    if (e < 36) atomicAdd(D + e, r == c ? v + lambda : v);
  }
  if (lane < 6) {
    double *bsch = sysacc + 36 * (size_t)P.n_blocks + 6 * (size_t)q;
    atomicAdd(bsch + lane, s);               // bschur
    bsch[6 * (size_t)P.n_fp + lane] = s;     // b_p (kept for computeScale)
  }
}

__global__ void __launch_bounds__(32 * kSchurWarps, 3) k_schur(const DeviceProblem P, int n_pose_ctas, int prefolded) {
  const Control *ctl = P.ctl;
  extern __shared__ double s_w_all[];  // kSchurWarps x kSchurRunPairs x 18
  __shared__ double s_dinv[kSchurWarps][kSchurRun][6];
  __shared__ double s_db[kSchurWarps][kSchurRun][3];
  __shared__ int s_pair0[kSchurWarps][kSchurRun];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // ---- static part (index structure only): runs while the previous kernel of the stream is still busy
  const bool pose_cta = (int)blockIdx.x < n_pose_ctas;
  const int u = ((int)blockIdx.x - n_pose_ctas) * kSchurWarps + warp;
  int s0 = 0, n = 0, k = 0, c0 = 0;
  if (pose_cta) {
    const int q = blockIdx.x * kSchurWarps + warp;
    if (q < P.n_fp) { prefetch_l1(P.q_part_ptr + q); prefetch_l1(P.col_diag + q); }
  } else if (u < P.n_units) {
    s0 = P.unit_slot[u]; n = P.unit_n[u]; k = P.unit_k[u]; c0 = P.unit_c0[u];
    if (lane < n) s_pair0[warp][lane] = P.slot_pair_ptr[s0 + lane];
    prefetch_l1((P.deterministic ? P.combo_pos : P.combo_blk) + P.unit_combo_ptr[u] + lane);
  }
  griddep_wait();
  griddep_launch();
  if (ctl->done) return;
  if (pose_cta) {
    const int q = blockIdx.x * kSchurWarps + warp;
    if (q < P.n_fp) schur_pose_warp(P, ctl, q, lane, prefolded != 0);
    return;
  }
  if (u >= P.n_units) return;  // whole warp; no block-wide barrier below
  const int lin = ctl->lin;
  const double *__restrict__ Wl = P.W[lin], *__restrict__ Hlll = P.Hll[lin], *__restrict__ bll = P.bl[lin];
  double *s_w = s_w_all + (size_t)warp * kSchurRunPairs * 18;
  const bool staged = n * k <= kSchurRunPairs;
  __syncwarp();
  if (staged) {
    // the run's W blocks (k consecutive 6x3 blocks per landmark) -> shared memory, asynchronously
    const int per_lm = 9 * k;  // 16-byte granules per landmark
    for (int gi = lane; gi < n * per_lm; gi += 32) {
      const int i = gi / per_lm, o = gi - i * per_lm;
      const unsigned sa = (unsigned)__cvta_generic_to_shared(s_w + 18 * (size_t)(i * k) + 2 * o);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa),
                   "l"(Wl + 18 * (size_t)s_pair0[warp][i] + 2 * o));
    }
    asm volatile("cp.async.commit_group;\n" ::);
  }
  const double lambda = ctl->lambda;
  if (lane < n) {
    const size_t sl = (size_t)(s0 + lane);
    double H[6], Di[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) H[i] = Hlll[6 * sl + i];
    H[0] += lambda; H[3] += lambda; H[5] += lambda;
    sym3_inverse(H, Di);
    const double b0 = bll[3 * sl], b1 = bll[3 * sl + 1], b2 = bll[3 * sl + 2];
#pragma unroll
    for (int i = 0; i < 6; ++i) s_dinv[warp][lane][i] = Di[i];
    s_db[warp][lane][0] = Di[0] * b0 + Di[1] * b1 + Di[2] * b2;
    s_db[warp][lane][1] = Di[1] * b0 + Di[3] * b1 + Di[4] * b2;
    s_db[warp][lane][2] = Di[2] * b0 + Di[4] * b1 + Di[5] * b2;
    if (c0 == 0) {
#pragma unroll
      for (int i = 0; i < 6; ++i) P.Dinv[6 * sl + i] = Di[i];
    }
  }
  asm volatile("cp.async.wait_all;\n" ::: "memory");
  __syncwarp();
  // cw block pairs of this unit, shared by G lane groups that split the run's landmarks
  const int ncomb = k * (k + 1) / 2;
  const int cw = ncomb - c0 < 32 ? ncomb - c0 : 32;
  int G = 32 / cw;
  G = G > 4 ? 4 : G;
  const int sub = lane / cw, ci = lane - sub * cw;
  const bool active = sub < G;
  const int idx = c0 + ci;
  // block pair (a <= b) of this lane, enumerated like combo_blk: a-major
  int a = 0, rem = idx;
  while (rem >= k - a) { rem -= k - a; ++a; }
  const int b = a + rem;
  const bool diag = a == b;
  double acc[36], accb[6];
#pragma unroll
  for (int i = 0; i < 36; ++i) acc[i] = 0.0;
#pragma unroll
  for (int i = 0; i < 6; ++i) accb[i] = 0.0;
  if (active) {
    for (int i = sub; i < n; i += G) {
      const double2 *sa, *sb;
      if (staged) {
        sa = reinterpret_cast<const double2 *>(s_w + 18 * (size_t)(i * k + a));
        sb = reinterpret_cast<const double2 *>(s_w + 18 * (size_t)(i * k + b));
      } else {
        const int p0 = s_pair0[warp][i];
        sa = reinterpret_cast<const double2 *>(Wl + 18 * (size_t)(p0 + a));
        sb = reinterpret_cast<const double2 *>(Wl + 18 * (size_t)(p0 + b));
      }
      // BD = W_a Dinv, two rows of W_a per step (three 16-byte loads); kept short-lived on purpose:
      // with all of W_a, W_b and BD live the kernel needs 222 registers, i.e. two CTAs per SM and
      // a second wave of CTAs; this way it fits three
      const double d0 = s_dinv[warp][i][0], d1 = s_dinv[warp][i][1], d2 = s_dinv[warp][i][2],
                   d3 = s_dinv[warp][i][3], d4 = s_dinv[warp][i][4], d5 = s_dinv[warp][i][5];
      const double e0 = s_db[warp][i][0], e1 = s_db[warp][i][1], e2 = s_db[warp][i][2];
      double BD[18];
#pragma unroll
      for (int t = 0; t < 3; ++t) {
        const double2 x0 = sa[3 * t], x1 = sa[3 * t + 1], x2 = sa[3 * t + 2];
        const double w[6] = {x0.x, x0.y, x1.x, x1.y, x2.x, x2.y};
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int r = 2 * t + h;
          const double w0 = w[3 * h], w1 = w[3 * h + 1], w2 = w[3 * h + 2];
          BD[3 * r + 0] = w0 * d0 + w1 * d1 + w2 * d2;
          BD[3 * r + 1] = w0 * d1 + w1 * d3 + w2 * d4;
          BD[3 * r + 2] = w0 * d2 + w1 * d4 + w2 * d5;
          if (diag) accb[r] += w0 * e0 + w1 * e1 + w2 * e2;
        }
      }
      // block (row q_b, col q_a) += W_b Dinv W_a^T, two rows of W_b per step
#pragma unroll
      for (int t = 0; t < 3; ++t) {
        const double2 x0 = sb[3 * t], x1 = sb[3 * t + 1], x2 = sb[3 * t + 2];
        const double w[6] = {x0.x, x0.y, x1.x, x1.y, x2.x, x2.y};
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int r = 2 * t + h;
#pragma unroll
          for (int c = 0; c < 6; ++c)
            acc[6 * r + c] += w[3 * h] * BD[3 * c] + w[3 * h + 1] * BD[3 * c + 1] + w[3 * h + 2] * BD[3 * c + 2];
        }
      }
    }
  }
  // the G partial blocks -> lane group 0 (fixed order, so the unit total is deterministic)
  for (int g = 1; g < G; ++g) {
#pragma unroll
    for (int i = 0; i < 36; ++i) {
      const double t = __shfl_down_sync(0xffffffffu, acc[i], g * cw);
      if (sub == 0) acc[i] += t;
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const double t = __shfl_down_sync(0xffffffffu, accb[i], g * cw);
      if (sub == 0) accb[i] += t;
    }
  }
  if (sub != 0) return;
  if (P.deterministic) {
    // the unit's totals go to its own place; k_schur_reduce adds the producers of every block in a fixed order
    const size_t cidx = (size_t)P.combo_pos[P.unit_combo_ptr[u] + ci];
    double2 *dst2 = reinterpret_cast<double2 *>(P.stage + 36 * cidx);
#pragma unroll
    for (int i = 0; i < 18; ++i) dst2[i] = make_double2(-acc[2 * i], -acc[2 * i + 1]);
    if (diag) {
      double2 *db2 = reinterpret_cast<double2 *>(P.stage_b + 6 * cidx);
#pragma unroll
      for (int i = 0; i < 3; ++i) db2[i] = make_double2(-accb[2 * i], -accb[2 * i + 1]);
    }
    return;
  }
  double *sysacc = schur_target(P, ctl);
  double *dst = sysacc + 36 * (size_t)P.combo_blk[P.unit_combo_ptr[u] + ci];
#pragma unroll
  for (int i = 0; i < 36; ++i) atomicAdd(dst + i, -acc[i]);
  if (diag) {
    double *bs = sysacc + 36 * (size_t)P.n_blocks + 6 * (size_t)P.pair_q[s_pair0[warp][0] + a];
#pragma unroll
    for (int r = 0; r < 6; ++r) atomicAdd(bs + r, -accb[r]);
  }
}

// k_schur_reduce - the deterministic accumulation of the reduced system: one CTA per factor block adds the totals
// the units of k_schur stored for it (contiguous in the staging buffer, in unit order: a list the host made) on top
// of Hpp + lambda I for a diagonal block (block_solver.hpp:334-335, 524-539); bschur = b_p - sum of the units'
// vectors (:397); blocks of the fill-in that no landmark touches become zero.  The eight warps take an eighth of
// the producers each (a diagonal block of a window has a hundred of them: every run of landmarks the pose sees)
// and the partial sums are combined as ((w0 + w1) + (w2 + w3)) + ((w4 + w5) + (w6 + w7)): a fixed order, no fp64
// atomics anywhere, so two runs give the same bits.
constexpr int kRedWarps = 8;
__global__ void __launch_bounds__(32 * kRedWarps) k_schur_reduce(const DeviceProblem P) {
  const Control *ctl = P.ctl;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.x;
  __shared__ double2 s_part[kRedWarps][24];  // per warp: 18 pieces of the block, 3 of the vector
  const int row = P.blk_row[b], col = P.blk_col[b], p0 = P.blk_prod_ptr[b], p1 = P.blk_prod_ptr[b + 1];
  griddep_wait();
  griddep_launch();
  if (ctl->done) return;
  const bool diag = row == col;
  // this warp's quarter of the producers; lanes 0..17 own one 16-byte piece of the block, lanes 24..26 of the vector
  const int per = (p1 - p0 + kRedWarps - 1) / kRedWarps;
  const int q0 = min(p1, p0 + warp * per), q1 = min(p1, q0 + per);
  double2 acc = make_double2(0.0, 0.0);
  if (lane < 18) {
    const double2 *src = reinterpret_cast<const double2 *>(P.stage) + 18 * (size_t)q0 + lane;
    // four loads in flight, added in producer order (the sum keeps its fixed order)
    int p = q0;
    for (; p + 4 <= q1; p += 4, src += 72) {
      const double2 v0 = __ldcg(src), v1 = __ldcg(src + 18), v2 = __ldcg(src + 36), v3 = __ldcg(src + 54);
      acc.x += v0.x; acc.y += v0.y; acc.x += v1.x; acc.y += v1.y; acc.x += v2.x; acc.y += v2.y; acc.x += v3.x; acc.y += v3.y;
    }
    for (; p < q1; ++p, src += 18) { const double2 v = __ldcg(src); acc.x += v.x; acc.y += v.y; }
  } else if (diag && lane >= 24 && lane < 27) {
    const double2 *src = reinterpret_cast<const double2 *>(P.stage_b) + 3 * (size_t)q0 + (lane - 24);
    for (int p = q0; p < q1; ++p, src += 3) { const double2 v = __ldcg(src); acc.x += v.x; acc.y += v.y; }
  }
  if (lane < 18) s_part[warp][lane] = acc;
  else if (lane >= 24 && lane < 27) s_part[warp][lane - 6] = acc;
  __syncthreads();
  if (warp != 0) return;
  const double lambda = ctl->rank == 0 ? ctl->lambda : 0.0;
  double *sysacc = schur_target(P, ctl);
  if (lane < 18) {
    double2 base = make_double2(0.0, 0.0);
    if (diag) {
      auto tri = [](int e) { const int r = e / 6, c = e - 6 * r; const int lo = r < c ? r : c, hi = r < c ? c : r; return 6 + lo * 6 - lo * (lo - 1) / 2 + (hi - lo); };
      const int e0 = 2 * lane, e1 = 2 * lane + 1;
      base.x = P.hpp_fold[27 * col + tri(e0)] + ((e0 / 6 == e0 % 6) ? lambda : 0.0);
      base.y = P.hpp_fold[27 * col + tri(e1)] + ((e1 / 6 == e1 % 6) ? lambda : 0.0);
    }
    const double2 a0 = s_part[0][lane], a1 = s_part[1][lane], a2 = s_part[2][lane], a3 = s_part[3][lane];
    const double2 a4 = s_part[4][lane], a5 = s_part[5][lane], a6 = s_part[6][lane], a7 = s_part[7][lane];
    double2 out;
    out.x = base.x + (((a0.x + a1.x) + (a2.x + a3.x)) + ((a4.x + a5.x) + (a6.x + a7.x)));
    out.y = base.y + (((a0.y + a1.y) + (a2.y + a3.y)) + ((a4.y + a5.y) + (a6.y + a7.y)));
    reinterpret_cast<double2 *>(sysacc + 36 * (size_t)b)[lane] = out;
  } else if (diag && lane >= 24 && lane < 27) {
    const int k = lane - 24;
    const double2 bp = make_double2(P.hpp_fold[27 * col + 2 * k], P.hpp_fold[27 * col + 2 * k + 1]);
    const double2 a0 = s_part[0][18 + k], a1 = s_part[1][18 + k], a2 = s_part[2][18 + k], a3 = s_part[3][18 + k];
    const double2 a4 = s_part[4][18 + k], a5 = s_part[5][18 + k], a6 = s_part[6][18 + k], a7 = s_part[7][18 + k];
    double2 *bsch = reinterpret_cast<double2 *>(sysacc + 36 * (size_t)P.n_blocks + 6 * (size_t)col);
    bsch[k] = make_double2(bp.x + (((a0.x + a1.x) + (a2.x + a3.x)) + ((a4.x + a5.x) + (a6.x + a7.x))),
                           bp.y + (((a0.y + a1.y) + (a2.y + a3.y)) + ((a4.y + a5.y) + (a6.y + a7.y))));  // bschur
    bsch[3 * (size_t)P.n_fp + k] = bp;                                                                         // b_p (kept for computeScale)
  }
}

// ---------------------------------------------------------------------------------------------
// k_reduced_solve — the reduced pose system S x_p = bschur: block-sparse left-looking Cholesky
// scheduled over the elimination-tree levels of the host's symbolic factorisation (natural or
// nested-dissection order), forward substitution fused into the same levels, backward
// substitution over the levels in reverse, then the pose part of SparseOptimizer::update
// (sparse_optimizer.cpp:433-446) and of computeScale (levenberg.cpp:168-175).  Replaces
// LinearSolverCSparse::solve (g2o/solvers/csparse/linear_solver_csparse.h:106-142) +
// cs_chol_workspace / cs_lsolve / cs_ltsolve (csparse_extension.cpp:35-122), incl. the
// "pivot <= 0 => the trial is rejected" rule (:115).
//
// One CTA.  Every 6x6 block of a column of the running level is one work item handled by one
// warp: lanes (g, r) = (lane / 6, lane % 6) split the item's update pairs five ways (g) and own
// one row (r) of the block.  The warp holding a column's diagonal item factors it across lanes
// 0..5 with shuffles, inverts the 6x6 triangle and publishes the inverse in shared memory; the
// warps holding the column's other items wait on a shared-memory flag, multiply by that inverse
// and store the final block.  One __syncthreads per level.

__device__ __forceinline__ double group_reduce(double v, int lane) {
  // lanes r + 6 g, g = 0..4 (lanes 30, 31 carry zeros): ((g0 + g3) + (g1 + g4)) + g2 -> lanes 0..5
  double t = __shfl_down_sync(0xffffffffu, v, 18);
  if (lane < 12) v += t;
  t = __shfl_down_sync(0xffffffffu, v, 6);
  const double u = __shfl_down_sync(0xffffffffu, v, 12);
  if (lane < 6) v = (v + t) + u;
  return v;
}

// the index segment of one level, resolved from its header (ssba_structure.cpp build_solver_program)
struct LevelSeg {
  int n_cols, n_rounds, n_pairs, n_pf, n_bpf;
  const int *cta_rptr;  // kSolveMaxCluster + 1: the rounds of CTA c are [cta_rptr[c], cta_rptr[c + 1])
  const int *col_j, *col_b0, *col_bptr, *brow, *round_type, *gt_dst, *gt_slot, *gt_pos, *gt_p0, *gt_p1, *gt_mask, *pa,
      *pb, *pf_blk, *pf_slot, *bpf_blk;
  __device__ __forceinline__ explicit LevelSeg(const int *seg) {
    n_cols = seg[0]; n_rounds = seg[1]; n_pairs = seg[2]; n_pf = seg[4]; n_bpf = seg[5];
    const int n_brows = seg[3];
    const int *p = seg + 8;
    cta_rptr = p; p += kSolveMaxCluster + 1;
    col_j = p; p += n_cols;
    col_b0 = p; p += n_cols;
    col_bptr = p; p += n_cols + 1;
    brow = p; p += n_brows;
    round_type = p; p += n_rounds;
    gt_dst = p; p += 5 * n_rounds;
    gt_slot = p; p += 5 * n_rounds;
    gt_pos = p; p += 5 * n_rounds;
    gt_p0 = p; p += 5 * n_rounds;
    gt_p1 = p; p += 5 * n_rounds;
    gt_mask = p; p += 5 * n_rounds;
    pa = p; p += n_pairs;
    pb = p; p += n_pairs;
    pf_blk = p; p += n_pf;
    pf_slot = p; p += n_pf;
    bpf_blk = p;
  }
};

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem_src));
}
// asynchronous global -> shared copy of one level segment (16-byte granules, LDGSTS)
__device__ __forceinline__ void stage_segment(int *dst, const int *src, int n_ints, int tid) {
  for (int i = 4 * tid; i < n_ints; i += 4 * kSolveThreads) cp_async16(dst + i, src + i);
}
// ---- bulk asynchronous copy (TMA, cp.async.bulk) completing on an mbarrier: one instruction for a
// whole level program.  (Measured: it does NOT pay for the 288-byte factor blocks, ~70 cycles per
// copy through the one TMA unit; those use 16-byte LDGSTS granules or plain loads.)
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, unsigned bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void stage_wait() {
  asm volatile("cp.async.commit_group;\n" ::);
  asm volatile("cp.async.wait_all;\n" ::: "memory");
}

// ---- thread-block cluster plumbing of k_reduced_solve: the columns of a level are dealt over the
// CTAs of the cluster (every SM brings its own shared-memory pipe, which is what bounds the
// update products).  All rounds of a column run in one CTA, so the inverse diagonal block and
// its flag stay local (measured: publishing it through DSMEM costs ~2.7 k cycles on the critical
// path of every level); the finished blocks and y are written into the block cache of EVERY CTA
// through distributed shared memory, off the critical path, so that all reads stay local.
__device__ __forceinline__ unsigned cluster_cta_rank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned dsmem_addr(const void *local_smem, unsigned cta) {
  unsigned ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"((unsigned)__cvta_generic_to_shared(local_smem)), "r"(cta));
  return ra;
}
__device__ __forceinline__ void dsmem_st2(unsigned ra, double a, double b) {
  asm volatile("st.shared::cluster.v2.f64 [%0], {%1, %2};" ::"r"(ra), "d"(a), "d"(b) : "memory");
}
__device__ __forceinline__ void dsmem_st1(unsigned ra, double a) {
  asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(ra), "d"(a) : "memory");
}
__device__ __forceinline__ void dsmem_st_int(unsigned ra, int v) {
  asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(ra), "r"(v) : "memory");
}
#ifdef SSBA_SOLVER_TRACE
__device__ long long g_solver_trace[4096];
#define TRACE(i) do { if (threadIdx.x == 0 && blockIdx.x == 0 && (i) < 4096) g_solver_trace[(i)] = clock64(); } while (0)
#define TRACE2(sgv, k) do { if (lane == 0 && ((sgv) == 6 || (sgv) == 18) && rd < 16) g_solver_trace[3000 + ((sgv) == 18 ? 512 : 0) + 16 * rd + (k)] = clock64(); } while (0)
// k_update: phase stamps (globaltimer, ns) of CTAs 0, 300 and 700 at entries 3800 + 16 * {0, 1, 2}
#define UTRACE(k) do { if (threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == 300 || blockIdx.x == 700)) g_solver_trace[3800 + 16 * (blockIdx.x == 0 ? 0 : blockIdx.x == 300 ? 1 : 2) + (k)] = gtime_early(); } while (0)
#else
#define TRACE(i) do { } while (0)
#define TRACE2(sgv, k) do { } while (0)
#define UTRACE(k) do { } while (0)
#endif

template <int kCluster>
__global__ void __launch_bounds__(kSolveThreads) k_reduced_solve(const DeviceProblem P, const SolverSmemLayout lay) {
  Control *ctl = P.ctl;
  if (ctl->done) return;  // uniform over the cluster
  const int cta = kCluster > 1 ? (int)cluster_cta_rank() : 0;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *s_linv = reinterpret_cast<double *>(smem_raw);                          // kSolveMaxCols x 36
  volatile int *s_flag = reinterpret_cast<volatile int *>(smem_raw + lay.off_flag);  // kSolveMaxCols
  double *xs = lay.x_in_smem ? reinterpret_cast<double *>(smem_raw + lay.off_x) : P.xp;  // y / x vector
  int *s_prog = reinterpret_cast<int *>(smem_raw + lay.off_prog);                 // 2 level segments
  double *s_slots = reinterpret_cast<double *>(smem_raw + lay.off_slots);         // factor cache
  const int staged = lay.staged;
  __shared__ int s_fail;
  __shared__ double red[kSolveThreads / 32];
  __shared__ __align__(8) unsigned long long s_bar[2];  // prefetch use u completes on s_bar[u & 1], parity (u >> 1) & 1
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = kSolveThreads / 32;
  const int g = lane / 6, r = lane - 6 * g;  // lanes 30, 31: g == 5, never active
  const int gl = 6 * g;                      // first lane of this lane's group
  double *L = P.sys;
  const double *bs = P.sys + 36 * (size_t)P.n_blocks;
  const double *bp = bs + 6 * (size_t)P.n_fp;
  const int seg_stride = (P.prog_max_seg + 3) & ~3;
  const int NSEG = P.n_segments;
  if (tid == 0) {
    s_fail = 0;
    mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < kSolveMaxCols) s_flag[tid] = 0;
  if (staged && NSEG > 0) stage_segment(s_prog, P.prog + P.prog_ptr[0], P.prog_ptr[1] - P.prog_ptr[0], tid);
  stage_wait();
  __syncthreads();
  if (kCluster > 1) cluster_barrier();  // flags / s_fail of every CTA initialised before anyone publishes
  // a block reference r >= 0 is a shared-memory slot, r < 0 is block -1-r in global memory
  auto deref = [&](int ref) -> const double * {
    return ref >= 0 ? s_slots + 36 * ref : L + 36 * (size_t)(-1 - ref);
  };

  // ---- numeric factorisation + forward substitution, segment by segment (0 = prologue)
  TRACE(0);
  for (int sg = 0; sg < NSEG; ++sg) {
    const int stamp = sg + 1;
    TRACE(1 + 3 * sg);
    const LevelSeg S(staged ? s_prog + (sg & 1) * seg_stride : P.prog + P.prog_ptr[sg]);
    // asynchronously: the next segment's program, and the initial values (Schur complement) of
    // the next level's blocks into their cache slots
    // Only the program is staged ahead: the initial value of a block (its Schur complement
    // entry) is read from global memory when its round starts and is not needed before the
    // round's update products are done, so that latency hides by itself.
    if (tid == kSolveThreads - 1) {
      const bool more = staged && sg + 1 < NSEG;
      const unsigned bytes = more ? 4u * (unsigned)(P.prog_ptr[sg + 2] - P.prog_ptr[sg + 1]) : 0u;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the buffer's last generic reads are two levels old
      mbar_arrive_expect_tx(&s_bar[sg & 1], bytes);
      if (more) bulk_g2s(s_prog + ((sg + 1) & 1) * seg_stride, P.prog + P.prog_ptr[sg + 1], bytes, &s_bar[sg & 1]);
    }
    for (int rd = S.cta_rptr[cta] + warp, rd1 = S.cta_rptr[cta + 1]; rd < rd1; rd += NW) {
      const int type = S.round_type[rd];
      const int kind = type & 3;
      const bool reduce = (type & 4) != 0;
      const int gt = 5 * rd + (g < 5 ? g : 0);
      bool active = g < 5 && S.gt_dst[gt] >= 0;
      const int dst = S.gt_dst[gt], slot = S.gt_slot[gt], pos = S.gt_pos[gt];
      const int p0 = active ? S.gt_p0[gt] : 0, p1 = active ? S.gt_p1[gt] : 0;
      double *linv = s_linv + 36 * pos;
      TRACE2(sg, 0);
      if (kind != 2) {
        double *D = L + 36 * (size_t)(active ? dst : 0);
        double *Ds = slot >= 0 ? s_slots + 36 * slot : D;
        double v[6];
#pragma unroll
        for (int c = 0; c < 6; ++c) v[c] = 0.0;
        if (active && (!reduce || g == 0)) {
          // from L2: another CTA's look-ahead round may have updated the block one level ago
          const double2 *d2 = reinterpret_cast<const double2 *>(D + 6 * r);
          const double2 t0 = __ldcg(d2), t1 = __ldcg(d2 + 1), t2 = __ldcg(d2 + 2);
          v[0] = t0.x; v[1] = t0.y; v[2] = t1.x; v[3] = t1.y; v[4] = t2.x; v[5] = t2.y;
        }
        double acc[6];
#pragma unroll
        for (int c = 0; c < 6; ++c) acc[c] = 0.0;
        TRACE2(sg, 5);
        for (int p = p0; p < p1; ++p) {
          const int ra = S.pa[p], rb = S.pb[p];
          if (ra >= 0 && rb >= 0) {
            // both blocks in the shared-memory cache (the usual case): explicit LDS instead of
            // generic loads
            const double2 *A2 = reinterpret_cast<const double2 *>(s_slots + 36 * ra + 6 * r);
            const double2 *B2 = reinterpret_cast<const double2 *>(s_slots + 36 * rb);
            const double2 a0 = A2[0], a1 = A2[1], a2 = A2[2];
#pragma unroll
            for (int c = 0; c < 6; ++c) {
              const double2 b0 = B2[3 * c], b1 = B2[3 * c + 1], b2 = B2[3 * c + 2];
              acc[c] += a0.x * b0.x + a0.y * b0.y + a1.x * b1.x + a1.y * b1.y + a2.x * b2.x + a2.y * b2.y;
            }
          } else {
            const double2 *A2 = reinterpret_cast<const double2 *>(deref(ra) + 6 * r);
            const double2 *B2 = reinterpret_cast<const double2 *>(deref(rb));
            const double2 a0 = A2[0], a1 = A2[1], a2 = A2[2];
#pragma unroll
            for (int c = 0; c < 6; ++c) {
              const double2 b0 = B2[3 * c], b1 = B2[3 * c + 1], b2 = B2[3 * c + 2];
              acc[c] += a0.x * b0.x + a0.y * b0.y + a1.x * b1.x + a1.y * b1.y + a2.x * b2.x + a2.y * b2.y;
            }
          }
        }
        TRACE2(sg, 1);
        if (reduce) {
#pragma unroll
          for (int c = 0; c < 6; ++c) acc[c] = group_reduce(acc[c], lane);  // totals on lanes 0..5
          active = active && g == 0;
        }
        if (active) {
#pragma unroll
          for (int c = 0; c < 6; ++c) v[c] -= acc[c];
        }
        if (kind == 3) {
          // ACC (look-ahead): the products of columns that finished two or more levels ago are
          // subtracted from the block of a column of the NEXT level ahead of time, in global
          // memory, where that block's own round will pick its value up
          if (active) {
            double2 *d2 = reinterpret_cast<double2 *>(D + 6 * r);
            d2[0] = make_double2(v[0], v[1]); d2[1] = make_double2(v[2], v[3]); d2[2] = make_double2(v[4], v[5]);
          }
        } else if (kind == 0) {
          // DIAG: the group's rows meet in shared memory, then the group's first lane factors the
          // 6x6 block and inverts the triangle serially in registers (measured on B200: ~2.5x
          // shorter than the same recurrences spread over six lanes with shuffles, whose
          // latency sits on the critical path of every level)
          if (active) {
#pragma unroll
            for (int c = 0; c < 6; ++c) linv[6 * r + c] = v[c];
          }
          __syncwarp();
          TRACE2(sg, 2);
          if (active && r == 0) {
            double a[36], invd[6];
#pragma unroll
            for (int i = 0; i < 6; ++i)
#pragma unroll
              for (int k = 0; k <= i; ++k) a[6 * i + k] = linv[6 * i + k];
            bool bad = false;
#pragma unroll
            for (int j = 0; j < 6; ++j) {
              double dj = a[7 * j];
              if (!(dj > 0.0)) { bad = true; dj = 1.0; }
              const double inv = rsqrt(dj);
              invd[j] = inv;
#pragma unroll
              for (int i = j + 1; i < 6; ++i) a[6 * i + j] *= inv;
#pragma unroll
              for (int i = j + 1; i < 6; ++i)
#pragma unroll
                for (int k = j + 1; k <= i; ++k) a[6 * i + k] -= a[6 * i + j] * a[6 * k + j];
            }
            // X = L^-1 (lower triangular), column by column
            double X[36];
#pragma unroll
            for (int i = 0; i < 36; ++i) X[i] = 0.0;
#pragma unroll
            for (int c = 0; c < 6; ++c) {
              X[7 * c] = invd[c];
#pragma unroll
              for (int i = c + 1; i < 6; ++i) {
                double sacc = 0.0;
#pragma unroll
                for (int k = c; k < i; ++k) sacc += a[6 * i + k] * X[6 * k + c];
                X[6 * i + c] = -invd[i] * sacc;
              }
            }
            double2 *l2 = reinterpret_cast<double2 *>(linv), *d2 = reinterpret_cast<double2 *>(D);
#pragma unroll
            for (int i = 0; i < 18; ++i) { const double2 t = make_double2(X[2 * i], X[2 * i + 1]); l2[i] = t; d2[i] = t; }
            if (bad) {
              if (kCluster == 1) s_fail = 1;
              else for (int cc = 0; cc < kCluster; ++cc) dsmem_st_int(dsmem_addr(&s_fail, (unsigned)cc), 1);
            }
          }
          TRACE2(sg, 3);
          // every round of a column runs in the CTA of its DIAG round: the inverse stays local
          __threadfence_block();
          __syncwarp();
          if (active && r == 0) s_flag[pos] = stamp;
          TRACE2(sg, 4);
        } else {
          // SUB: X = B L_jj^-T once the column's inverse diagonal block is out
          while (!__all_sync(0xffffffffu, !active || s_flag[pos] == stamp)) { }
          __threadfence_block();
          if (active) {
            double x[6];
#pragma unroll
            for (int c = 0; c < 6; ++c) {
              double sacc = 0.0;
#pragma unroll
              for (int k = 0; k <= c; ++k) sacc += v[k] * linv[6 * c + k];
              x[c] = sacc;
            }
            double2 *d2 = reinterpret_cast<double2 *>(D + 6 * r);
            d2[0] = make_double2(x[0], x[1]); d2[1] = make_double2(x[2], x[3]); d2[2] = make_double2(x[4], x[5]);
            if (slot >= 0) {
              if (kCluster == 1) {
                double2 *s2 = reinterpret_cast<double2 *>(Ds + 6 * r);
                s2[0] = make_double2(x[0], x[1]); s2[1] = make_double2(x[2], x[3]); s2[2] = make_double2(x[4], x[5]);
              } else {
                const int mask = S.gt_mask[gt];  // the CTAs that read this block at a later level
#pragma unroll
                for (int cc = 0; cc < kCluster; ++cc) {
                  if (!((mask >> cc) & 1)) continue;
                  const unsigned ra = dsmem_addr(Ds + 6 * r, (unsigned)cc);
                  dsmem_st2(ra, x[0], x[1]); dsmem_st2(ra + 16, x[2], x[3]); dsmem_st2(ra + 32, x[4], x[5]);
                }
              }
            }
          }
        }
      } else {
        // VEC: forward substitution y_j = L_jj^-1 (bschur_j - sum_k L(j,k) y_k)
        const int j = active ? S.col_j[pos] : 0;
        double acc = 0.0;
        for (int p = p0; p < p1; ++p) {
          const double2 *B2 = reinterpret_cast<const double2 *>(deref(S.pa[p]) + 6 * r);
          const double *yk = xs + 6 * S.pb[p];
          const double2 b0 = B2[0], b1 = B2[1], b2 = B2[2];
          acc += b0.x * yk[0] + b0.y * yk[1] + b1.x * yk[2] + b1.y * yk[3] + b2.x * yk[4] + b2.y * yk[5];
        }
        if (reduce) { acc = group_reduce(acc, lane); active = active && g == 0; }
        const double sv = active ? bs[6 * j + r] - acc : 0.0;
        while (!__all_sync(0xffffffffu, !active || s_flag[pos] == stamp)) { }
        __threadfence_block();
        double y = 0.0;
#pragma unroll
        for (int c = 0; c < 6; ++c) {
          const double sc = __shfl_sync(0xffffffffu, sv, gl + c);
          if (active && c <= r) y += linv[6 * r + c] * sc;
        }
        if (active) {
          if (kCluster == 1 || !lay.x_in_smem) xs[6 * j + r] = y;
          else {
#pragma unroll
            for (int cc = 0; cc < kCluster; ++cc) dsmem_st1(dsmem_addr(xs + 6 * j + r, (unsigned)cc), y);
          }
        }
      }
      TRACE2(sg, 6);
    }
    TRACE(2 + 3 * sg);
    mbar_wait(&s_bar[sg & 1], (unsigned)(sg >> 1) & 1u);
    if (kCluster == 1) __syncthreads(); else cluster_barrier();
    TRACE(3 + 3 * sg);
  }
  // the backward pass and the pose update are latency-bound chains: CTA 0 runs them alone (the
  // factor is complete in global memory, and nobody touches another CTA's memory from here on)
  if (cta != 0) return;
  const bool fail = s_fail != 0;

  if (!fail) {
    // ---- backward: x_j = L_jj^-T (y_j - sum_{i > j} L(i,j)^T x_i), segments in reverse.  The
    // program of the last segment is still in its buffer.  While level sg runs, the blocks of
    // level sg-1 are copied into one half of the (now free) block cache.
    const int half = lay.n_slots / 2;
    for (int sg = NSEG - 1; sg >= 1; --sg) {
      const LevelSeg S(staged ? s_prog + (sg & 1) * seg_stride : P.prog + P.prog_ptr[sg]);
      const int use = NSEG + (NSEG - 1 - sg);  // prefetch uses continue the numbering of the forward pass
      if (tid == kSolveThreads - 1) {
        const bool more = staged && sg > 1;
        const unsigned bytes = more ? 4u * (unsigned)(P.prog_ptr[sg] - P.prog_ptr[sg - 1]) : 0u;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive_expect_tx(&s_bar[use & 1], bytes);
        if (more) bulk_g2s(s_prog + ((sg - 1) & 1) * seg_stride, P.prog + P.prog_ptr[sg - 1], bytes, &s_bar[use & 1]);
      }
      if (warp >= NW / 2) {  // the columns of a level keep the first warps busy
        const int st = tid - 32 * (NW / 2);
        const int nb = S.n_bpf < half ? S.n_bpf : half;
        double *dstb = s_slots + 36 * (size_t)(((sg - 1) & 1) * half);
        for (int i = st; i < 18 * nb; i += 32 * (NW / 2)) {
          const int e = i / 18, o = i - 18 * e;
          cp_async16(dstb + 36 * e + 2 * o, L + 36 * (size_t)S.bpf_blk[e] + 2 * o);
        }
      }
      const bool cached = sg != NSEG - 1;
      const double *srcb = s_slots + 36 * (size_t)((sg & 1) * half);
      // one warp per column: its sub-diagonal blocks are split over the five lane groups
      for (int t = warp; t < S.n_cols; t += NW) {
        const int j = S.col_j[t];
        const int b0 = S.col_b0[t];
        const int nb = S.col_bptr[t + 1] - S.col_bptr[t];
        const int *rows = S.brow + S.col_bptr[t];
        const int l0 = S.col_bptr[t] + t;  // level-local index of the diagonal block
        double acc = 0.0;
        if (g < 5) {
          for (int k = g; k < nb; k += 5) {
            const int li = l0 + 1 + k;
            const double *B = (cached && li < half) ? srcb + 36 * li : L + 36 * (size_t)(b0 + 1 + k);
            const double *xi = xs + 6 * rows[k];
#pragma unroll
            for (int c = 0; c < 6; ++c) acc += B[6 * c + r] * xi[c];
          }
        }
        acc = group_reduce(acc, lane);  // totals on lanes 0..5
        const double sv = lane < 6 ? xs[6 * j + r] - acc : 0.0;
        const double *Li = (cached && l0 < half) ? srcb + 36 * l0 : L + 36 * (size_t)b0;  // inverse diagonal block
        double xv = 0.0;
#pragma unroll
        for (int c = 0; c < 6; ++c) {
          const double sc = __shfl_sync(0xffffffffu, sv, c);
          if (lane < 6 && c >= r) xv += Li[6 * c + r] * sc;
        }
        if (lane < 6) xs[6 * j + r] = xv;
      }
      stage_wait();
      mbar_wait(&s_bar[use & 1], (unsigned)(use >> 1) & 1u);
      __syncthreads();
      TRACE(3 * NSEG + 1 + (NSEG - sg));
    }
  } else {
    for (int i = tid; i < 6 * P.n_fp; i += kSolveThreads) xs[i] = 0.0;
    __syncthreads();
  }

  // ---- pose part of computeScale and of update(): T <- exp(x) T into the trial buffer
  const int cur = ctl->cur;
  const double lambda = ctl->lambda;
  double sc = 0.0;
  for (int i = tid; i < 6 * P.n_fp; i += kSolveThreads) {
    const double xi = xs[i];
    sc += xi * (lambda * xi + bp[i]);
    if (lay.x_in_smem) P.xp[i] = xi;
  }
  sc = block_sum<kSolveThreads>(sc, red);
  for (int q = tid; q < P.n_fp; q += kSolveThreads) {
    const int kv = P.pose_of_q[q];
    double T[7], d[6], out[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) T[i] = P.pose[cur][7 * kv + i];
#pragma unroll
    for (int i = 0; i < 6; ++i) d[i] = xs[6 * q + i];
    pose_oplus(T, d, out);
#pragma unroll
    for (int i = 0; i < 7; ++i) P.pose[cur ^ 1][7 * kv + i] = fail ? T[i] : out[i];
  }
  if (tid == 0) {
    ctl->scale_pose_part[0] = sc;
    ctl->chol_fail = fail ? 1 : 0;
  }
  TRACE(4 * NSEG + 2);
}

// ---------------------------------------------------------------------------------------------
// control_step — the accept/reject law of OptimizationAlgorithmLevenberg::solve
// (levenberg.cpp:119-149) and the iteration bookkeeping of SparseOptimizer::optimize
// (sparse_optimizer.cpp:386-426), on the device so that no host round trip sits between trials.
// Runs on one thread once scal[0..2] = chi(current), chi(trial), landmark part of computeScale.
__device__ void control_step(const DeviceProblem &P, bool trial_linearized) {
  Control *c = P.ctl;
  c->trial_seq++;
  const double currentChi = P.scal[0];
  double tempChi = P.scal[1];
  if (c->outer_iter == 0 && c->qmax == 0) c->chi2_initial = currentChi;
  const bool fail = c->chol_fail != 0;
  if (fail) { tempChi = DBL_MAX; c->cholesky_failures++; }
  c->chol_fail = 0;  // k_tree_solve only ever raises it
  double scale_pose = 0.0;
  for (int i = 0; i < kTreeMaxCluster; ++i) scale_pose += c->scale_pose_part[i];  // CTA order: deterministic
  double scale = P.scal[2] + scale_pose;
  scale += 1e-3;
  double rho = (currentChi - tempChi) / scale;
  if (fail) rho = -1.0;  // a failed factorisation always rejects the step (levenberg.cpp:120-121)
  bool broke = false, accepted = false;
  if (rho > 0 && isfinite(tempChi)) {
    const double t = 2 * rho - 1;
    double alpha = 1. - t * t * t;
    alpha = fmin(alpha, c->good_upper);
    const double scaleFactor = fmax(c->good_lower, alpha);
    c->lambda *= scaleFactor;
    c->ni = 2;
    c->cur ^= 1;  // discardTop(): the trial buffer becomes the estimate
    if (trial_linearized) c->lin ^= 1;  // ... and k_update's linearisation of it the current one
    accepted = true;
  } else {
    c->lambda *= c->ni;
    c->ni *= 2;  // pop(): the estimate buffer is simply kept
    if (!isfinite(c->lambda)) broke = true;
  }
  if (!broke) c->qmax++;
  c->n_trials++;
  const bool stop = *reinterpret_cast<volatile int *>(&c->force_stop) != 0;  // terminate(), levenberg.cpp:145
  c->rho = rho;
  c->temp_chi = tempChi;
  c->current_chi = accepted ? tempChi : currentChi;
  const bool again = !broke && rho < 0 && c->qmax < c->max_trials && !stop;
  if (again) {
    c->need_linearize = 0;
    c->need_fold = 0;
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
  // the next iteration starts from a linearisation: the one k_update made of the accepted trial, or k_linearize's
  c->need_linearize = (accepted && trial_linearized) ? 0 : 1;
  c->need_fold = (accepted && trial_linearized) ? 1 : 0;
  c->lin_valid = (accepted && trial_linearized) ? 1 : 0;
  // k_update linearises: the host does not enqueue k_linearize with every slot.  An iteration that ends without an
  // accepted step AND without Terminate (rho is NaN) needs one: pause (every kernel returns on done != 0) until the
  // host has seen it and enqueues a linearising slot.
  if (trial_linearized && c->need_linearize) c->done = 2;
  if (result != SSBA_SOLVER_OK || c->outer_iter >= c->max_iters || stop) c->done = 1;  // sparse_optimizer.cpp:388
}

// partial sums of the linearize / update CTAs -> scal[0..2], folded in a fixed order by one CTA
template <int NT>
__device__ __forceinline__ void fold_partials(const DeviceProblem &P, double *red) {
  // eight loads in flight per thread, added in index order
  auto strided_sum = [](const double *__restrict__ part, int n) {
    double acc = 0.0;
    for (int base = threadIdx.x; base < n; base += 8 * NT) {
      double v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = base + j * NT < n ? __ldcg(part + base + j * NT) : 0.0;  // written by other CTAs of this launch
#pragma unroll
      for (int j = 0; j < 8; ++j) acc += v[j];
    }
    return acc;
  };
  // chi2 of the current state: k_linearize's partial sums when it ran in this slot, else the value the last decision
  // left (the accepted trial's chi2, or the unchanged one of a re-trial); several ranks add their parts up, so only
  // rank 0 contributes the kept value
  const Control *ctl = P.ctl;
  double a = ctl->need_linearize ? strided_sum(P.chi_cur_part, P.n_lin_blocks) : 0.0;
  double b = strided_sum(P.chi_new_part, P.n_upd_blocks);
  double c = strided_sum(P.scale_part, P.n_upd_blocks);
  a = block_sum<NT>(a, red);
  b = block_sum<NT>(b, red);
  c = block_sum<NT>(c, red);
  if (threadIdx.x == 0) {
    if (!ctl->need_linearize) a = ctl->rank == 0 ? ctl->current_chi : 0.0;
    P.scal[0] = a; P.scal[1] = b; P.scal[2] = c;
  }
}

// ---------------------------------------------------------------------------------------------
// k_update — landmark back-substitution x_l = Dinv (b_l - W^T x_p) (block_solver.hpp:420-442),
// p <- p + x_l (g2otypes.hpp:54-59), the landmark part of computeScale, and the trial
// computeActiveErrors + activeRobustChi2 (levenberg.cpp:116-117).  Same decomposition as
// k_linearize: one thread per pair (W^T x_p, then the pair's trial residuals), the landmark's
// first thread folds the pairs and moves the point.
__device__ __forceinline__ double pair_trial_chi(const DeviceProblem &P, int a, const double *pose_new, const double *p) {
  const int kv = P.pair_vertex[a];
  double T[7];
#pragma unroll
  for (int i = 0; i < 7; ++i) T[i] = pose_new[7 * kv + i];
  double chi = 0.0;
  const int e1 = P.pair_edge_ptr[a + 1];
  for (int e = P.pair_edge_ptr[a]; e < e1; ++e) {
    EdgeTerms t;
    edge_error(P.cams.K, P.cams.ext[P.e_cam[e]], T, p, P.e_uv[2 * e], P.e_uv[2 * e + 1], t.e0, t.e1);
    load_edge_weighting(P, P.e_info, P.e_delta, e, t);
    chi += t.rho0;
  }
  return chi;
}

__device__ __forceinline__ void pair_wtx(const DeviceProblem &P, const double *__restrict__ Wl, int a, int q, double *c3) {
  c3[0] = c3[1] = c3[2] = 0.0;
  if (q < 0) return;
  const double2 *src = reinterpret_cast<const double2 *>(Wl + 18 * (size_t)a);
  double Wv[18];
#pragma unroll
  for (int i = 0; i < 9; ++i) { const double2 t = src[i]; Wv[2 * i] = t.x; Wv[2 * i + 1] = t.y; }
#pragma unroll
  for (int r = 0; r < 6; ++r) {
    const double xr = P.xp[6 * q + r];
    c3[0] += Wv[3 * r] * xr; c3[1] += Wv[3 * r + 1] * xr; c3[2] += Wv[3 * r + 2] * xr;
  }
}

//
// The reduced system is dead once k_reduced_solve has run: every CTA zeroes a slice of it for the
// atomic accumulation of the next trial's k_schur.  kFusedControl (single GPU): the CTA that
// finishes last folds the partial sums and takes the accept/reject decision (control_step), so
// no separate launch sits between two trials.
// kFusedLin (closed-form Jacobians): the trial residuals are not just summed - the trial state is LINEARISED
// into the other set of buffers (W, Hll, b_l, Hpp partials), exactly as k_linearize would do at the start of the
// next iteration if the trial is accepted.  control_step then flips Control::lin together with Control::cur and the
// next slot's k_linearize has nothing to do (three launches per accepted iteration); a rejected trial keeps the
// old buffers, like the reference keeps its linearisation across the trials of one iteration.
template <bool kFusedControl, bool kFusedLin>
__global__ void __launch_bounds__(kLinThreads, kFusedLin ? SSBA_LIN_MINB : 4) k_update(const DeviceProblem P) {
  Control *ctl = P.ctl;
  __shared__ LinSmem sm;
  __shared__ int s_last;
  double *red = sm.red;
  // ---- static part (index structure only): runs while the reduced solve is still busy on its few SMs; the index
  // lines this CTA is going to walk are pulled into L1 now
  const int s0 = P.lchunk_slot[blockIdx.x], s1 = P.lchunk_slot[blockIdx.x + 1];
  const int a0 = P.slot_pair_ptr[s0], a1 = P.slot_pair_ptr[s1];
  const int tid = threadIdx.x;
  UTRACE(0);
  PairRec my = {0, -1, 0, 0, 0, 0, 0, 0};  // this thread's pair (chunks of <= 128 pairs), in registers before the wait
  {
    const int a = a0 + tid;
    if (a < a1 && a1 - a0 <= kLinThreads) {
      my = load_pair_rec(P, a);
      prefetch_l1(P.e_uv + 2 * (size_t)my.e0); prefetch_l1(P.e_cam + my.e0);
      prefetch_l1(P.pair_rec + a);  // linearize_chunk loads it again
    }
    if (s0 + tid < s1) { prefetch_l1(P.slot_pair_ptr + s0 + tid); prefetch_l1(P.slot_vertex + s0 + tid); prefetch_l1(P.slot_free + s0 + tid); }
    if (tid == 0) { prefetch_l1(P.lchunk_lp_ptr + blockIdx.x); }
  }
  UTRACE(1);
  griddep_wait();
  griddep_launch();
  UTRACE(2);
  if (ctl->done) return;
  __shared__ __align__(8) unsigned long long s_bar;
  extern __shared__ __align__(16) double s_pose[];
  const unsigned pb = kFusedLin ? P.pose_stage : 0u;
  if (pb) ek_stage_begin(&s_bar, s_pose, P.pose[ctl->cur ^ 1], pb);  // the trial poses: needed by the third phase only
  if (!P.deterministic) {
    const size_t nz = 36 * (size_t)P.n_blocks + 6 * (size_t)P.n_fp;  // blocks and bschur; b_p is overwritten
    // peer-memory exchange: the partial the NEXT trial accumulates into is the one the peers read
    // one trial ago; their flag_scal of that trial (seen by the last k_control) says they are done
    double *z = P.use_p2p ? reinterpret_cast<double *>(P.peer[ctl->rank] + sizeof(PeerHeader)) + (size_t)(ctl->trial_seq & 1) * P.sys_doubles
                          : P.sys;
    for (size_t i = (size_t)blockIdx.x * kLinThreads + threadIdx.x; i < nz; i += (size_t)gridDim.x * kLinThreads) z[i] = 0.0;
  }
  double (*s_part)[9] = sm.part;          // [a][0..2]: W^T x_p of the pair
  double (*s_pnew)[27] = sm.hp;           // [slot - s0][0..2]: the landmark's trial position (plain path)
  const int cur = ctl->cur, lin = ctl->lin;
  const bool fail = ctl->chol_fail != 0;
  const double lambda = ctl->lambda;
  const double *__restrict__ pose_new = P.pose[cur ^ 1];
  const double *__restrict__ Wl = P.W[lin], *__restrict__ bll = P.bl[lin];
  const bool small = a1 - a0 <= kLinThreads;
  double chi = 0.0, scale = 0.0;
  // ---- W^T x_p per pair
  double c3[3] = {0.0, 0.0, 0.0};
  if (small) {
    if (tid < a1 - a0 && my.lfree) pair_wtx(P, Wl, a0 + tid, my.q, c3);
    s_part[tid][0] = c3[0]; s_part[tid][1] = c3[1]; s_part[tid][2] = c3[2];
  } else if (P.slot_free[s0]) {
    for (int a = a0 + tid; a < a1; a += kLinThreads) {
      double t3[3];
      pair_wtx(P, Wl, a, P.pair_q[a], t3);
      c3[0] += t3[0]; c3[1] += t3[1]; c3[2] += t3[2];
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) { const double v = block_sum<kLinThreads>(c3[i], red); if (tid == 0) s_part[0][i] = v; }
  }
  __syncthreads();
  UTRACE(3);
  // ---- the landmark's first thread: x_l, scale, new position
  {
    const int sl = s0 + tid;
    if (sl < s1) {
      const int pv = P.slot_vertex[sl];
      double p[3] = {P.point[cur][3 * pv], P.point[cur][3 * pv + 1], P.point[cur][3 * pv + 2]};
      if (P.slot_free[sl]) {
        const double b0 = bll[3 * (size_t)sl], b1 = bll[3 * (size_t)sl + 1], b2 = bll[3 * (size_t)sl + 2];
        double c0 = b0, c1 = b1, c2 = b2;
        if (small) {
          for (int a = P.slot_pair_ptr[sl] - a0; a < P.slot_pair_ptr[sl + 1] - a0; ++a) {
            c0 -= s_part[a][0]; c1 -= s_part[a][1]; c2 -= s_part[a][2];
          }
        } else {
          c0 -= s_part[0][0]; c1 -= s_part[0][1]; c2 -= s_part[0][2];
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
      if (!kFusedLin) { s_pnew[tid][0] = p[0]; s_pnew[tid][1] = p[1]; s_pnew[tid][2] = p[2]; }
    }
  }
  __syncthreads();  // the trial positions of this chunk's landmarks are in place (global memory, written by this CTA)
  UTRACE(4);
  // ---- trial residuals per pair (and, fused, the whole linearisation of the trial state)
  if (kFusedLin) {
    double mx;
    if (pb) ek_stage_wait(&s_bar);  // (two __syncthreads since ek_stage_begin)
    linearize_chunk<SSBA_JACOBIAN_ANALYTIC>(P, lin ^ 1, pb ? s_pose : pose_new, P.point[cur ^ 1], sm, chi, mx);
    (void)mx;  // lambda_0 uses the first linearisation only (k_linearize)
  } else if (small) {
    if (tid < a1 - a0) {
      const int a = a0 + tid;
      chi = pair_trial_chi(P, a, pose_new, s_pnew[P.pair_slot[a] - s0]);
    }
  } else {
    for (int a = a0 + tid; a < a1; a += kLinThreads) chi += pair_trial_chi(P, a, pose_new, s_pnew[0]);
  }
  UTRACE(5);
  const double s_ = block_sum<kLinThreads>(chi, red);
  const double sc = block_sum<kLinThreads>(scale, red);
  if (tid == 0) { P.chi_new_part[blockIdx.x] = s_; P.scale_part[blockIdx.x] = sc; }
  UTRACE(6);
  if (kFusedControl) {
    if (tid == 0) {
      __threadfence();
      s_last = atomicAdd(&ctl->ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;  // block-uniform
    __threadfence();
    fold_partials<kLinThreads>(P, red);
    if (tid == 0) { ctl->ticket = 0; control_step(P, kFusedLin); }
  }
}

// several GPUs: partial sums -> scal[0..2] (then all-reduced), and the decision in k_control
__global__ void __launch_bounds__(256) k_reduce_partials(const DeviceProblem P) {
  if (P.ctl->done) return;
  __shared__ double red[8];
  fold_partials<256>(P, red);
}

__global__ void k_control(const DeviceProblem P, int trial_linearized) {
  if (P.ctl->done) return;
  control_step(P, trial_linearized != 0);
}

// ---- several GPUs: the two exchanges of a trial through NVLink peer memory instead of NCCL calls.
// Every rank owns an exchange buffer that all ranks have mapped (CUDA IPC).  Flags and the small
// partial sums are pushed into the peers' headers (remote stores, release at system scope), so a
// rank polls only its own memory; the partial reduced systems are pulled from the peers once their
// flag is in.  Every rank adds the parts in rank order, so all ranks hold bitwise the same sums.  Parts are double-buffered by trial
// parity: a buffer is rewritten two trials later, after flags that can only have been published
// once every peer was done reading it.  A peer that does not show up within about a minute sets
// comm_timeout (reported as SSBA_ERR_NCCL) instead of hanging the GPU.
__device__ __forceinline__ long long ld_acquire_sys(const long long *p) {
  long long v;
  asm volatile("ld.acquire.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(long long *p, long long v) {
  asm volatile("st.release.sys.global.s64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ bool wait_peer_flag(const long long *flag, long long seq) {
  for (long long spin = 0; spin < (1ll << 27); ++spin) {  // about a minute: a peer may still be building its structure
    if (ld_acquire_sys(flag) >= seq) return true;
    __nanosleep(40);
  }
  return false;
}

__device__ __forceinline__ void st_f64_sys(double *p, double v) {
  asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}

__device__ __forceinline__ long long gtime() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

constexpr int kXchgThreads = 256;
__global__ void __launch_bounds__(kXchgThreads) k_exchange_sys(const DeviceProblem P) {
  Control *ctl = P.ctl;
  if (ctl->done) return;
  if (blockIdx.x == 0 && threadIdx.x == 0) ctl->dbg[0] = gtime();
  const long long seq = ctl->trial_seq + 1;  // the running trial
  const int world = ctl->world, rank = ctl->rank, par = (int)(seq & 1);
  __shared__ int s_ok;
  if (threadIdx.x < 32) {
    // k_schur of this trial is complete (stream order; its sums sit in this GPU's L2, which is where
    // the peers read them): lane r of CTA 0 tells rank r (release, system scope); every CTA then
    // waits for the peers' words in this rank's own header - no CTA waits for another one of this launch.
    const int r = threadIdx.x;
    bool ok = true;
    if (r < world) {
      if (blockIdx.x == 0) st_release_sys(&reinterpret_cast<PeerHeader *>(P.peer[r])->flag_sys_from[rank], seq);
      if (r != rank) ok = wait_peer_flag(&reinterpret_cast<const PeerHeader *>(P.peer[rank])->flag_sys_from[r], seq);
    }
    ok = __all_sync(0xffffffffu, ok);
    if (threadIdx.x == 0) s_ok = ok ? 1 : 0;
  }
  __syncthreads();
  if (blockIdx.x == 0 && threadIdx.x == 0) ctl->dbg[1] = gtime();
  if (!s_ok) { if (threadIdx.x == 0) ctl->comm_timeout = 1; }
  const size_t n2 = P.sys_doubles / 2;  // sys_doubles is even: 36 n_blocks + 12 n_fp
  double2 *out = reinterpret_cast<double2 *>(P.sys);
  for (size_t i = (size_t)blockIdx.x * kXchgThreads + threadIdx.x; i < n2; i += (size_t)gridDim.x * kXchgThreads) {
    double2 acc = make_double2(0.0, 0.0);
    for (int r = 0; r < world; ++r) {  // rank order: every rank gets bitwise the same sum
      const double2 v = __ldcg(reinterpret_cast<const double2 *>(P.peer[r] + sizeof(PeerHeader)) + (size_t)par * n2 + i);
      acc.x += v.x; acc.y += v.y;
    }
    out[i] = acc;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) ctl->dbg[2] = gtime();
}

__global__ void __launch_bounds__(256) k_control_p2p(const DeviceProblem P, int trial_linearized) {
  Control *ctl = P.ctl;
  if (ctl->done) return;
  if (threadIdx.x == 0) ctl->dbg[3] = gtime();
  __shared__ double red[8];
  fold_partials<256>(P, red);  // this rank's sums -> scal[0..2]
  __syncthreads();
  if (threadIdx.x == 0) ctl->dbg[4] = gtime();
  if (threadIdx.x >= 32) return;
  const long long seq = ctl->trial_seq + 1;
  const int world = ctl->world, rank = ctl->rank, par = (int)(seq & 1);
  const int r = threadIdx.x;
  double a = 0.0, b = 0.0, c = 0.0;
  bool ok = true;
  if (r < world) {
    // push this rank's three sums and the flag into rank r's header, then wait for rank r's in ours
    PeerHeader *dst = reinterpret_cast<PeerHeader *>(P.peer[r]);
    for (int i = 0; i < 3; ++i) st_f64_sys(&dst->scal_from[rank][par][i], P.scal[i]);
    st_release_sys(&dst->flag_scal_from[rank], seq);
    const PeerHeader *mine = reinterpret_cast<const PeerHeader *>(P.peer[rank]);
    ok = wait_peer_flag(&mine->flag_scal_from[r], seq);
    a = __ldcg(&mine->scal_from[r][par][0]); b = __ldcg(&mine->scal_from[r][par][1]); c = __ldcg(&mine->scal_from[r][par][2]);
  }
  ok = __all_sync(0xffffffffu, ok);
  double sa = 0.0, sb = 0.0, sc = 0.0;
  for (int k = 0; k < world; ++k) {  // rank order
    sa += __shfl_sync(0xffffffffu, a, k); sb += __shfl_sync(0xffffffffu, b, k); sc += __shfl_sync(0xffffffffu, c, k);
  }
  if (threadIdx.x != 0) return;
  ctl->dbg[5] = gtime();
  if (!ok) ctl->comm_timeout = 1;
  P.scal[0] = sa; P.scal[1] = sb; P.scal[2] = sc;
  control_step(P, trial_linearized != 0);
  ctl->dbg[6] = gtime();
}

// ---------------------------------------------------------------------------------------------
// final read-outs at the current estimate: activeChi2 / activeRobustChi2
// (sparse_optimizer.cpp:92-116), the outlier count of backend.cpp:180-197, per-edge errors.
__global__ void __launch_bounds__(kReadoutThreads) k_final_chi2(const DeviceProblem P, double threshold,
                                                                int write_errors) {
  __shared__ double red[kReadoutThreads / 32];
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
        if (write_errors == 1) {
          const int o = P.e_orig[e];
          P.err_out[2 * (size_t)o] = t.e0; P.err_out[2 * (size_t)o + 1] = t.e1;
        } else if (write_errors == 2) {
          P.mask_out[P.e_orig[e]] = t.chi > threshold ? 1 : 0;
        }
      }
    }
  }
  const double a = block_sum<kReadoutThreads>(plain, red);
  const double b = block_sum<kReadoutThreads>(robust, red);
  const double c = block_sum<kReadoutThreads>(nout, red);
  const double d = block_sum<kReadoutThreads>(nin, red);
  if (threadIdx.x == 0) {
    // the partial arrays of the LM loop are free at this point
    P.chi_cur_part[blockIdx.x] = a; P.maxdiag_part[blockIdx.x] = b;
    P.chi_new_part[blockIdx.x] = c; P.scale_part[blockIdx.x] = d;
  }
}

__global__ void __launch_bounds__(256) k_final_reduce(const DeviceProblem P) {
  __shared__ double red[8];
  double a = 0.0, b = 0.0, c = 0.0, d = 0.0;
  for (int i = threadIdx.x; i < P.n_fin_blocks; i += 256) {
    a += P.chi_cur_part[i]; b += P.maxdiag_part[i]; c += P.chi_new_part[i]; d += P.scale_part[i];
  }
  a = block_sum<256>(a, red); b = block_sum<256>(b, red);
  c = block_sum<256>(c, red); d = block_sum<256>(d, red);
  if (threadIdx.x == 0) { P.chi_out[0] = a; P.chi_out[1] = b; P.chi_out[2] = c; P.chi_out[3] = d; }
}

// The measurement values arrive in the caller's edge order (one plain copy of the caller's arrays); this puts them
// into the structure's landmark-major order - the host structure build never touches them.
__global__ void __launch_bounds__(256) k_gather_edge_values(const DeviceProblem P, const double *__restrict__ raw_uv,
                                                            const double *__restrict__ raw_info, const double *__restrict__ raw_delta,
                                                            double *e_uv, double *e_info, double *e_delta) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.n_edges) return;
  const size_t o = (size_t)P.e_orig[i];
  reinterpret_cast<double2 *>(e_uv)[i] = reinterpret_cast<const double2 *>(raw_uv)[o];
  if (e_info) { e_info[3 * (size_t)i] = raw_info[3 * o]; e_info[3 * (size_t)i + 1] = raw_info[3 * o + 1]; e_info[3 * (size_t)i + 2] = raw_info[3 * o + 2]; }
  if (e_delta) e_delta[i] = raw_delta[o];
}

__global__ void __launch_bounds__(256) k_pack_pairs(const DeviceProblem P, PairRec *out) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= P.n_pairs) return;
  const int sl = P.pair_slot[a], e0 = P.pair_edge_ptr[a];
  PairRec r;
  r.pose_row = P.pair_vertex[a]; r.q = P.pair_q[a]; r.slot = sl; r.e0 = e0;
  r.n_edges = P.pair_edge_ptr[a + 1] - e0; r.point_row = P.slot_vertex[sl]; r.lfree = P.slot_free[sl] ? 1 : 0; r.pad = 0;
  out[a] = r;
}

__global__ void k_gather_points(const DeviceProblem P) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 3 * P.n_points) return;
  P.gather[i] = P.owner_mask[i / 3] ? P.point[P.ctl->cur][i] : 0.0;
}

// ---------------------------------------------------------------------------------------------
// Pose-graph optimisation (LoopClosing::PoseGraphOptimization, src/ssvio/loopclosing.cpp:458-532):
// EdgePoseGraph errors e = log(M^-1 T0 T1^-1) (g2otypes.hpp:169-176) with the numeric Jacobians the
// reference ships (base_binary_edge.hpp:144-212: central differences, delta = 1e-9, through
// VertexPose::oplusImpl), identity information, no robust kernel; H and b assembled with fp64
// atomics into the block pattern of the factor; the SAME level-scheduled solver (k_reduced_solve)
// and the same control law (control_step) as the bundle adjustment.  Graphs are small (one vertex
// per key-frame), so these kernels are plain: one thread per edge / per scalar.
struct PoseGraphArgs {
  int n_edges;
  const int32_t *ev0, *ev1;   // pose rows
  const int32_t *eq0, *eq1;   // block columns (q) or -1 for a fixed key-frame
  const int32_t *eblk;        // off-diagonal block (row max(q0,q1), col min) or -1
  const double *minv;         // n_edges x 7: measurement^-1
  double *sysH;               // H and b of the current linearisation, layout of DeviceProblem::sys
  // deterministic assembly: every edge stores its three 6x6 products and two 6-vectors (120 doubles); every block
  // / right-hand side of H sums its producers in edge order (lists made by the host): no atomics
  double *stage;              // n_edges x 120: [J0^T J0 | J1^T J1 | Jrow^T Jcol | -J0^T e | -J1^T e]
  const int32_t *prod_ptr;    // n_blocks + n_fp + 1: producers of every block, then of every right-hand side
  const int32_t *prod;        // offsets into `stage` (in doubles)
};

// One WARP per edge: the 24 perturbed evaluations of the central differences (2 vertices x 6 directions x +-delta)
// run on 24 lanes at once instead of one after the other in one thread (the reference's order of operations inside
// every evaluation is kept, so the Jacobian entries are the same numbers); the J^T J / J^T e products are then
// spread over the lanes.  (One thread per edge: 85 us per launch on 400 key-frames, the largest part of the whole
// pose-graph optimisation.)
constexpr int kPgWarps = 4;
__global__ void __launch_bounds__(32 * kPgWarps) k_pg_linearize(const DeviceProblem P, const PoseGraphArgs A) {
  const Control *ctl = P.ctl;
  if (ctl->done || !ctl->need_linearize) return;
  __shared__ double s_J[kPgWarps][2][36];  // per warp: J0, J1 row-major 6x6: J[6 * k + d] = d e_k / d delta_d
  __shared__ double s_err[kPgWarps][6];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int e = blockIdx.x * kPgWarps + warp;
  if (e >= A.n_edges) return;  // whole warp
  const double *pose = P.pose[ctl->cur];
  double T0[7], T1[7], Mi[7];
#pragma unroll
  for (int i = 0; i < 7; ++i) { T0[i] = pose[7 * A.ev0[e] + i]; T1[i] = pose[7 * A.ev1[e] + i]; Mi[i] = A.minv[7 * e + i]; }
  const int q0 = A.eq0[e], q1 = A.eq1[e];
  const double delta = 1e-9, scalar = 1 / (2 * delta);
  // lane = 12 * side + 2 * d + (0: +delta, 1: -delta); lanes 24..31: the unperturbed error (lane 24 keeps it)
  const int side = lane / 12, d = (lane % 12) >> 1, minus = lane & 1;
  double ev[6];
  if (lane < 24) {
    double add[6] = {0, 0, 0, 0, 0, 0}, Tp[7];
    add[d] = minus ? -delta : delta;
    if (side == 0) { pose_oplus(T0, add, Tp); pose_graph_error(Mi, Tp, T1, ev); }
    else { pose_oplus(T1, add, Tp); pose_graph_error(Mi, T0, Tp, ev); }
  } else {
    pose_graph_error(Mi, T0, T1, ev);
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const double em = __shfl_down_sync(0xffffffffu, ev[k], 1);  // the -delta evaluation sits on the next lane
    if (lane < 24 && !minus) s_J[warp][side][6 * k + d] = scalar * (ev[k] - em);
    if (lane == 24) s_err[warp][k] = ev[k];
  }
  __syncwarp();
  // constructQuadraticForm (base_binary_edge.hpp:61-134) with Omega = I and no robust kernel, into this edge's staging
  // record (k_pg_assemble sums the records of every block in edge order): 120 outputs over 32 lanes
  double *st = A.stage + 120 * (size_t)e;
  const double *J0 = s_J[warp][0], *J1 = s_J[warp][1], *err = s_err[warp];
  for (int o = lane; o < 120; o += 32) {
    double v = 0.0;
    if (o < 108) {
      const int part = o / 36, rc = o - 36 * part, r = rc / 6, c = rc - 6 * r;
      const double *Ja, *Jb;
      if (part == 0) { if (q0 < 0) continue; Ja = J0; Jb = J0; }
      else if (part == 1) { if (q1 < 0) continue; Ja = J1; Jb = J1; }
      else { if (q0 < 0 || q1 < 0) continue; Ja = q0 > q1 ? J0 : J1; Jb = q0 > q1 ? J1 : J0; }  // block (row max(q0, q1), col min): Jrow^T Jcol
#pragma unroll
      for (int k = 0; k < 6; ++k) v += Ja[6 * k + r] * Jb[6 * k + c];
    } else {
      const int sd = (o - 108) / 6, r = (o - 108) - 6 * sd;
      if ((sd == 0 ? q0 : q1) < 0) continue;
      const double *J = sd == 0 ? J0 : J1;
#pragma unroll
      for (int k = 0; k < 6; ++k) v -= J[6 * k + r] * err[k];
    }
    st[o] = v;
  }
}

// H and b of the linearisation: every scalar of every block (and right-hand side) = the sum of its producers' staged
// values, in edge order -> sysH (kept for the re-trials of the iteration)
__global__ void __launch_bounds__(256) k_pg_assemble(const DeviceProblem P, const PoseGraphArgs A) {
  const Control *ctl = P.ctl;
  if (ctl->done || !ctl->need_linearize) return;
  const size_t nb = 36 * (size_t)P.n_blocks, nv = 6 * (size_t)P.n_fp;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nb + nv; i += (size_t)gridDim.x * blockDim.x) {
    int list, o;
    if (i < nb) { list = (int)(i / 36); o = (int)(i - 36 * (size_t)list); }
    else { const size_t v = i - nb; list = P.n_blocks + (int)(v / 6); o = (int)(v % 6); }
    double acc = 0.0;
    for (int p = A.prod_ptr[list]; p < A.prod_ptr[list + 1]; ++p) acc += A.stage[A.prod[p] + o];
    A.sysH[i] = acc;
  }
}

// lambda_0 = tau * max |H_jj| (levenberg.cpp:152-166) at the first iteration
__global__ void __launch_bounds__(256) k_pg_lambda_init(const DeviceProblem P, const PoseGraphArgs A) {
  Control *ctl = P.ctl;
  if (ctl->done || !ctl->first_iteration) return;
  __shared__ double red[8];
  double m = 0.0;
  for (int i = threadIdx.x; i < 6 * P.n_fp; i += 256) m = fmax(m, fabs(A.sysH[36 * (size_t)P.col_diag[i / 6] + 7 * (i % 6)]));
  m = block_max<256>(m, red);
  if (threadIdx.x == 0) {
    ctl->maxdiag = m;
    ctl->lambda = ctl->user_lambda > 0 ? ctl->user_lambda : ctl->tau * m;
    ctl->ni = 2.0;
    ctl->first_iteration = 0;
  }
}

// the trial's system: H + lambda I (setLambda, block_solver.hpp:524-548), b twice (bschur, b_p)
__global__ void k_pg_prepare(const DeviceProblem P, const PoseGraphArgs A) {
  const Control *ctl = P.ctl;
  if (ctl->done) return;
  const double lambda = ctl->lambda;
  const size_t nb = 36 * (size_t)P.n_blocks, nv = 6 * (size_t)P.n_fp;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nb + nv; i += (size_t)gridDim.x * blockDim.x) {
    double v = A.sysH[i];
    if (i < nb) {
      const int b = (int)(i / 36), o = (int)(i - 36 * (size_t)b);
      if (P.blk_row[b] == P.blk_col[b] && o % 7 == 0) v += lambda;
      P.sys[i] = v;
    } else {
      P.sys[i] = v;
      P.sys[i + nv] = v;
    }
  }
}

// chi2 of the current and of the trial poses (activeRobustChi2 without a kernel = sum e^T e), then the
// accept / reject decision
__global__ void __launch_bounds__(256) k_pg_chi_control(const DeviceProblem P, const PoseGraphArgs A, int control) {
  Control *ctl = P.ctl;
  if (control && ctl->done) return;
  __shared__ double red[8];
  const int cur = ctl->cur;
  double a = 0.0, b = 0.0;
  for (int e = threadIdx.x; e < A.n_edges; e += 256) {
    double T0[7], T1[7], Mi[7], err[6];
#pragma unroll
    for (int i = 0; i < 7; ++i) Mi[i] = A.minv[7 * e + i];
#pragma unroll
    for (int w = 0; w < 2; ++w) {
      const double *pose = P.pose[cur ^ w];
#pragma unroll
      for (int i = 0; i < 7; ++i) { T0[i] = pose[7 * A.ev0[e] + i]; T1[i] = pose[7 * A.ev1[e] + i]; }
      pose_graph_error(Mi, T0, T1, err);
      double c2 = 0.0;
#pragma unroll
      for (int k = 0; k < 6; ++k) c2 += err[k] * err[k];
      if (w == 0) a += c2; else b += c2;
    }
  }
  a = block_sum<256>(a, red);
  b = block_sum<256>(b, red);
  if (threadIdx.x != 0) return;
  P.scal[0] = a; P.scal[1] = b; P.scal[2] = 0.0;
  if (control) control_step(P, false);
}

inline int div_up(long long a, int b) { return (int)((a + b - 1) / b); }

}  // namespace

// ------------------------------------------------------------------------------------------------

int kernels_per_linearize() { return 1; }

#ifdef SSBA_SOLVER_TRACE
extern "C" int ssba_debug_solver_trace(long long *out, int n) {
  cudaDeviceSynchronize();
  return (int)cudaMemcpyFromSymbol(out, g_solver_trace, sizeof(long long) * (n < 4096 ? n : 4096));
}
#endif

// function attributes are per device: `setup` runs once per device of this process, and a second thread
// cannot launch before the first one has finished setting them (handles are used from several threads)
template <class F>
void once_per_device(std::atomic<unsigned long long> &done, F &&setup) {
  static std::mutex mu;
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned long long bit = 1ull << (dev & 63);
  if (done.load(std::memory_order_acquire) & bit) return;
  std::lock_guard<std::mutex> lk(mu);
  if (done.load(std::memory_order_acquire) & bit) return;
  setup();
  done.fetch_or(bit, std::memory_order_release);
}

void launch_linearize(const DeviceProblem &P, cudaStream_t st) {
  if (P.n_lin_blocks <= 0) return;
  static std::atomic<unsigned long long> seen{0};
  once_per_device(seen, [] {
    cudaFuncSetAttribute(k_linearize<SSBA_JACOBIAN_NUMERIC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPoseStageMaxBytes);
    cudaFuncSetAttribute(k_linearize<SSBA_JACOBIAN_ANALYTIC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPoseStageMaxBytes);
  });
  const size_t dyn = P.pose_stage;
  if (P.jacobian_mode == SSBA_JACOBIAN_NUMERIC) k_linearize<SSBA_JACOBIAN_NUMERIC><<<P.n_lin_blocks, kLinThreads, dyn, st>>>(P);
  else k_linearize<SSBA_JACOBIAN_ANALYTIC><<<P.n_lin_blocks, kLinThreads, dyn, st>>>(P);
}

void launch_fold(const DeviceProblem &P, cudaStream_t st) {
  if (P.n_fp > 0) k_fold<<<div_up(P.n_fp, 4), 128, 0, st>>>(P);
}

void launch_maxdiag(const DeviceProblem &P, cudaStream_t st) { k_maxdiag<<<1, 1024, 0, st>>>(P); }

void launch_lambda_init(const DeviceProblem &P, cudaStream_t st) { k_lambda_init<<<1, 1, 0, st>>>(P); }

void launch_schur(const DeviceProblem &P, bool prefolded, cudaStream_t st) {
  static std::atomic<unsigned long long> seen{0};
  constexpr size_t kDyn = (size_t)kSchurWarps * kSchurRunPairs * 18 * sizeof(double);
  once_per_device(seen, [] { cudaFuncSetAttribute(k_schur, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDyn); });
  const int n_pose = div_up(P.n_fp, kSchurWarps), n_unit = div_up(P.n_units, kSchurWarps);
  if (n_pose + n_unit > 0) launch_maybe_pdl(k_schur, dim3(n_pose + n_unit), dim3(32 * kSchurWarps), kDyn, st, P.pdl != 0, P, n_pose, prefolded ? 1 : 0);
  if (P.deterministic && P.n_blocks > 0) launch_maybe_pdl(k_schur_reduce, dim3(P.n_blocks), dim3(32 * kRedWarps), 0, st, P.pdl != 0, P);
}

// The largest cluster (8, 4, 2 or 1 CTAs of kSolveThreads threads with the full dynamic shared memory)
// this device can co-schedule for k_reduced_solve; the host program builder is capped by it.
int max_solver_cluster() {
  cudaFuncSetAttribute(k_reduced_solve<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSolveMaxDynSmem);
  cudaFuncSetAttribute(k_reduced_solve<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSolveMaxDynSmem);
  cudaFuncSetAttribute(k_reduced_solve<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSolveMaxDynSmem);
  for (int c = 8; c >= 2; c /= 2) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(c); cfg.blockDim = dim3(kSolveThreads); cfg.dynamicSmemBytes = kSolveMaxDynSmem;
    cudaLaunchAttribute attr{};
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = c; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr; cfg.numAttrs = 1;
    int n = 0;
    cudaError_t e = c == 8 ? cudaOccupancyMaxActiveClusters(&n, k_reduced_solve<8>, &cfg)
                  : c == 4 ? cudaOccupancyMaxActiveClusters(&n, k_reduced_solve<4>, &cfg)
                           : cudaOccupancyMaxActiveClusters(&n, k_reduced_solve<2>, &cfg);
    if (e == cudaSuccess && n >= 1) return c;
    cudaGetLastError();
  }
  return 1;
}

void launch_reduced_solve(const DeviceProblem &P, cudaStream_t st) {
  if (P.tree.C > 0) { launch_tree_solve(P, st); return; }
  static std::atomic<unsigned long long> seen{0};
  once_per_device(seen, [] {
    cudaFuncSetAttribute(k_reduced_solve<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSolveMaxDynSmem);
    cudaFuncSetAttribute(k_reduced_solve<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSolveMaxDynSmem);
    cudaFuncSetAttribute(k_reduced_solve<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSolveMaxDynSmem);
    cudaFuncSetAttribute(k_reduced_solve<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSolveMaxDynSmem);
  });
  const SolverSmemLayout lay = solver_smem_layout(P.n_fp, P.prog_max_seg);
  const int c = P.solve_cluster;  // the program was built for this many CTAs
  if (c == 1) { k_reduced_solve<1><<<1, kSolveThreads, lay.bytes, st>>>(P, lay); return; }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(c); cfg.blockDim = dim3(kSolveThreads); cfg.dynamicSmemBytes = lay.bytes; cfg.stream = st;
  cudaLaunchAttribute attr{};
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = c; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
  cfg.attrs = &attr; cfg.numAttrs = 1;
  if (c == 2) cudaLaunchKernelEx(&cfg, k_reduced_solve<2>, P, lay);
  else if (c == 4) cudaLaunchKernelEx(&cfg, k_reduced_solve<4>, P, lay);
  else cudaLaunchKernelEx(&cfg, k_reduced_solve<8>, P, lay);
}

// k_update linearises the trial state itself (closed-form Jacobians; SSBA_FUSE_LIN=0 keeps the plain kernel)
bool update_linearizes(const DeviceProblem &P) {
  static const bool on = [] { const char *e = std::getenv("SSBA_FUSE_LIN"); return !(e && std::atoi(e) == 0); }();
  return on && P.jacobian_mode == SSBA_JACOBIAN_ANALYTIC;
}

void launch_update(const DeviceProblem &P, bool fused_control, cudaStream_t st) {
  if (P.n_upd_blocks <= 0) return;
  const bool lin = update_linearizes(P);
  static std::atomic<unsigned long long> seen{0};
  once_per_device(seen, [] {
    cudaFuncSetAttribute(k_update<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPoseStageMaxBytes);
    cudaFuncSetAttribute(k_update<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPoseStageMaxBytes);
  });
  const size_t dyn = lin ? P.pose_stage : 0;
  if (fused_control) {
    if (lin) launch_maybe_pdl(k_update<true, true>, dim3(P.n_upd_blocks), dim3(kLinThreads), dyn, st, P.pdl != 0, P);
    else launch_maybe_pdl(k_update<true, false>, dim3(P.n_upd_blocks), dim3(kLinThreads), 0, st, P.pdl != 0, P);
  } else {
    if (lin) launch_maybe_pdl(k_update<false, true>, dim3(P.n_upd_blocks), dim3(kLinThreads), dyn, st, P.pdl != 0, P);
    else launch_maybe_pdl(k_update<false, false>, dim3(P.n_upd_blocks), dim3(kLinThreads), 0, st, P.pdl != 0, P);
  }
}

void launch_reduce_partials(const DeviceProblem &P, cudaStream_t st) {
  k_reduce_partials<<<1, 256, 0, st>>>(P);
}

void launch_control(const DeviceProblem &P, cudaStream_t st) { k_control<<<1, 1, 0, st>>>(P, update_linearizes(P) ? 1 : 0); }

void launch_exchange_sys(const DeviceProblem &P, cudaStream_t st) {
  // remote loads over NVLink take microseconds: enough CTAs that every thread has one or two 16-byte pieces and the
  // whole pull is one round of loads in flight (cfg5 on 8 GPUs: 64 CTAs took 5 rounds, 119 us per trial with the
  // control exchange; two CTAs per SM bring it to a single round)
  const int n = (int)std::min<size_t>(2 * 148, (P.sys_doubles / 2 + kXchgThreads - 1) / kXchgThreads);
  k_exchange_sys<<<n > 0 ? n : 1, kXchgThreads, 0, st>>>(P);
}

void launch_control_p2p(const DeviceProblem &P, cudaStream_t st) { k_control_p2p<<<1, 256, 0, st>>>(P, update_linearizes(P) ? 1 : 0); }

void launch_final_chi2(const DeviceProblem &P, double threshold, cudaStream_t st) {
  k_final_chi2<<<P.n_fin_blocks, kReadoutThreads, 0, st>>>(P, threshold, 0);
  k_final_reduce<<<1, 256, 0, st>>>(P);
}

void launch_gather_edge_values(const DeviceProblem &P, const double *raw_uv, const double *raw_info, const double *raw_delta, cudaStream_t st) {
  if (P.n_edges <= 0) return;
  k_gather_edge_values<<<div_up(P.n_edges, 256), 256, 0, st>>>(P, raw_uv, raw_info, raw_delta, const_cast<double *>(P.e_uv),
                                                                 const_cast<double *>(P.e_info), const_cast<double *>(P.e_delta));
}

void launch_pack_pairs(const DeviceProblem &P, cudaStream_t st) {
  if (P.n_pairs > 0) k_pack_pairs<<<div_up(P.n_pairs, 256), 256, 0, st>>>(P, const_cast<PairRec *>(P.pair_rec));
}

void launch_gather_points(const DeviceProblem &P, cudaStream_t st) {
  if (P.n_points > 0) k_gather_points<<<div_up(3LL * P.n_points, 256), 256, 0, st>>>(P);
}

void launch_edge_errors(const DeviceProblem &P, cudaStream_t st) {
  k_final_chi2<<<P.n_fin_blocks, kReadoutThreads, 0, st>>>(P, 0.0, 1);
}

void launch_outlier_mask(const DeviceProblem &P, double threshold, cudaStream_t st) {
  k_final_chi2<<<P.n_fin_blocks, kReadoutThreads, 0, st>>>(P, threshold, 2);
  k_final_reduce<<<1, 256, 0, st>>>(P);
}

void launch_pose_graph_slot(const DeviceProblem &P, int n_edges, const int32_t *ev0, const int32_t *ev1, const int32_t *eq0,
                            const int32_t *eq1, const int32_t *eblk, const double *minv, double *sysH, double *stage,
                            const int32_t *prod_ptr, const int32_t *prod, bool first, cudaStream_t st) {
  PoseGraphArgs A{n_edges, ev0, ev1, eq0, eq1, eblk, minv, sysH, stage, prod_ptr, prod};
  const int nz = div_up((long long)P.sys_doubles, 256);
  if (n_edges > 0) k_pg_linearize<<<div_up(n_edges, kPgWarps), 32 * kPgWarps, 0, st>>>(P, A);
  k_pg_assemble<<<nz, 256, 0, st>>>(P, A);
  if (first) k_pg_lambda_init<<<1, 256, 0, st>>>(P, A);
  k_pg_prepare<<<nz, 256, 0, st>>>(P, A);
  launch_reduced_solve(P, st);
  k_pg_chi_control<<<1, 256, 0, st>>>(P, A, 1);
}

void launch_pose_graph_chi(const DeviceProblem &P, int n_edges, const int32_t *ev0, const int32_t *ev1, const double *minv,
                           cudaStream_t st) {
  PoseGraphArgs A{n_edges, ev0, ev1, nullptr, nullptr, nullptr, minv, nullptr, nullptr, nullptr, nullptr};
  k_pg_chi_control<<<1, 256, 0, st>>>(P, A, 0);
}

}  // namespace ssba
