// ssba_tree_solve.cu — k_tree_solve: the reduced pose system S x_p = bschur on a thread-block cluster,
// one set of elimination subtrees per CTA, entirely out of shared memory (program: ssba_tree_program.hpp).
// Replaces LinearSolverCSparse::solve (g2o/solvers/csparse/linear_solver_csparse.h:106-142) +
// cs_chol_workspace / cs_lsolve / cs_ltsolve (g2o/solvers/csparse/csparse_extension.cpp:35-122), incl.
// "pivot <= 0 => the trial is rejected" (:115), and applies the pose part of SparseOptimizer::update
// (g2o/core/sparse_optimizer.cpp:433-446) and of computeScale (optimization_algorithm_levenberg.cpp:168-175).
//
// Numerics: block LDL^T with closed-form 6x6 inverses (ssba_tree_program.hpp).
// Data movement: the CTA's slice of the reduced system (its factor blocks, right-hand sides, b_p) and its
// whole program arrive with four bulk asynchronous copies (TMA, cp.async.bulk) on one mbarrier; from then
// on every operand is a shared-memory load.  Two cluster barriers per solve (contributions up, top solution
// down); one block barrier per forward step (plus a named barrier of the look-ahead warps), one per backward step.
#include <atomic>

#include "ssba_device.hpp"
#include "ssba_block_inverse.cuh"

namespace ssba {

namespace {

__device__ __forceinline__ unsigned ts_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ts_mbar_init(unsigned long long *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ts_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void ts_mbar_arrive_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ts_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ts_mbar_wait(unsigned long long *bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(ts_smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void ts_bulk_g2s(void *smem_dst, const void *gmem_src, unsigned bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(ts_smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(ts_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ unsigned ts_cluster_ctarank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void ts_cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// lanes r + 6 g, g = 0..4 (lanes 30, 31 carry zeros): ((g0 + g3) + (g1 + g4)) + g2 -> lanes 0..5
__device__ __forceinline__ double ts_group_reduce(double v, int lane) {
  double t = __shfl_down_sync(0xffffffffu, v, 18);
  if (lane < 12) v += t;
  t = __shfl_down_sync(0xffffffffu, v, 6);
  const double u = __shfl_down_sync(0xffffffffu, v, 12);
  if (lane < 6) v = (v + t) + u;
  return v;
}

// row r of  dest -= sum over the item's pairs of A B^T  (A: the item's rows, 6 doubles each, of a scaled block Y
// or a vector w; B: the unscaled copy of a 6x6 block).  `pw` = the item's first pair word (the caller may have
// fetched it ahead).  Returns the updated row in `d` as well (the diagonal items go on in registers).
__device__ __forceinline__ void ts_product_rows(double *pool, const int *prog, int w0, int p0, unsigned pw, int r, double *d) {
  const int np = (int)((unsigned)w0 >> 20);
  double acc[6];
#pragma unroll
  for (int c = 0; c < 6; ++c) acc[c] = 0.0;
  double2 *D = reinterpret_cast<double2 *>(pool + (w0 & 0xffff) + 6 * r);
  const double2 d0 = D[0], d1 = D[1], d2 = D[2];
  for (int p = 0; p < np; ++p) {
    const double2 *A2 = reinterpret_cast<const double2 *>(pool + (pw & 0xffffu) + 6 * r);
    const double2 *B2 = reinterpret_cast<const double2 *>(pool + (pw >> 16));
    if (p + 1 < np) pw = (unsigned)prog[p0 + p + 1];
    const double2 a0 = A2[0], a1 = A2[1], a2 = A2[2];
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      const double2 b0 = B2[3 * c], b1 = B2[3 * c + 1], b2 = B2[3 * c + 2];
      acc[c] += a0.x * b0.x + a0.y * b0.y + a1.x * b1.x + a1.y * b1.y + a2.x * b2.x + a2.y * b2.y;
    }
  }
  d[0] = d0.x - acc[0]; d[1] = d0.y - acc[1]; d[2] = d1.x - acc[2]; d[3] = d1.y - acc[3]; d[4] = d2.x - acc[4]; d[5] = d2.y - acc[5];
  if (np > 0) { D[0] = make_double2(d[0], d[1]); D[1] = make_double2(d[2], d[3]); D[2] = make_double2(d[4], d[5]); }
}

template <int NT>
__device__ __forceinline__ double ts_block_sum(double v, double *sm /* NT / 32 */) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
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

#ifdef SSBA_SOLVER_TRACE
__device__ long long g_tree_trace[kTreeMaxCluster][256];
__device__ __forceinline__ long long ts_gtime() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
// entries 0..15: phase boundaries (globaltimer, ns); 16..: clock64 after interval 1 / interval 2 of every forward step, then per backward step
#define TS_TRACE_G(i) do { if (threadIdx.x == 0) g_tree_trace[cta][(i)] = ts_gtime(); } while (0)
#define TS_TRACE_C(i) do { if (threadIdx.x == 0 && (i) < 256) g_tree_trace[cta][(i)] = clock64(); } while (0)
// fine trace of warp 0 inside the intervals of every step: 8 stamps per step from entry 128 on (steps 0..15)
#define TS_TRACE_F(s, k) do { if (threadIdx.x == 0 && ((s) == 4 || (s) == 5)) g_tree_trace[cta][200 + 8 * ((s) - 4) + (k)] = clock64(); } while (0)
// per warp, steps 4 and 5: five stamps (diagonal warp: step start, scaled its blocks, products done, inverse done, barrier
// passed; look-ahead warp: step start, panel done, named barrier passed, look-ahead done, barrier passed) for warps 0..5
#define TS_TRACE_W(s, k) do { if ((threadIdx.x & 31) == 0 && ((s) == 4 || (s) == 5) && (threadIdx.x >> 5) < 6) g_tree_trace[cta][128 + 32 * ((s) - 4) + 5 * (threadIdx.x >> 5) + (k)] = clock64(); } while (0)
#else
#define TS_TRACE_F(s, k) do { } while (0)
#define TS_TRACE_W(s, k) do { } while (0)
#define TS_TRACE_G(i) do { } while (0)
#define TS_TRACE_C(i) do { } while (0)
#endif

template <bool kCluster>
__global__ void __launch_bounds__(kTreeThreads, 1) k_tree_solve(const DeviceProblem P) {
  Control *ctl = P.ctl;
  const TreeDev &T = P.tree;
  const int cta = kCluster ? (int)ts_cluster_ctarank() : 0;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane / 6, r = lane - 6 * g;  // lanes 30, 31: g == 5, never active
  extern __shared__ __align__(128) unsigned char ts_smem[];
  const int np = T.pool_doubles[cta];
  const int nwords = T.prog_ptr[cta + 1] - T.prog_ptr[cta];
  double *pool = reinterpret_cast<double *>(ts_smem);
  int *prog = reinterpret_cast<int *>(ts_smem + 8 * (size_t)((np + 1) & ~1));
  unsigned char *misc = reinterpret_cast<unsigned char *>(prog + nwords);  // nwords is a multiple of 4
  unsigned long long *bar = reinterpret_cast<unsigned long long *>(misc);
  int *s_fail = reinterpret_cast<int *>(misc + 12);
  double *red = reinterpret_cast<double *>(misc + 16);  // kTreeWarps doubles
  const int nblk = T.n_own_blocks[cta], ncols = T.n_own_cols[cta], q0 = T.q0[cta];
  const int V0 = 36 * nblk, BP0 = V0 + 6 * ncols;
  const double *bs = P.sys + 36 * (size_t)P.n_blocks;

  TS_TRACE_G(0);
  if (tid == 0) {
    ts_mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    *s_fail = 0;
  }
  __syncthreads();
  // the program is static: its copy starts while the previous kernel (k_schur) may still be running (programmatic
  // dependent launch); the reduced system itself only after griddep_wait()
  const unsigned bytes_blk = 288u * (unsigned)nblk, bytes_vec = 48u * (unsigned)ncols, bytes_prog = 4u * (unsigned)nwords;
  if (tid == 0) {
    ts_mbar_arrive_expect_tx(bar, bytes_blk + 2u * bytes_vec + bytes_prog);
    if (bytes_prog) ts_bulk_g2s(prog, T.prog + T.prog_ptr[cta], bytes_prog, bar);
  }
  griddep_wait();
  griddep_launch();
  if (ctl->done) {  // uniform over the cluster; the program copy in flight must land before the CTA may exit
    if (tid == 0) {
      if (bytes_blk) ts_bulk_g2s(pool, P.sys + 36 * (size_t)T.b0[cta], bytes_blk, bar);
      if (bytes_vec) { ts_bulk_g2s(pool + V0, bs + 6 * (size_t)q0, bytes_vec, bar); ts_bulk_g2s(pool + BP0, bs + 6 * (size_t)(P.n_fp + q0), bytes_vec, bar); }
    }
    ts_mbar_wait(bar, 0);
    return;
  }
  if (tid == 0) {
    if (bytes_blk) ts_bulk_g2s(pool, P.sys + 36 * (size_t)T.b0[cta], bytes_blk, bar);
    if (bytes_vec) {
      ts_bulk_g2s(pool + V0, bs + 6 * (size_t)q0, bytes_vec, bar);                  // bschur
      ts_bulk_g2s(pool + BP0, bs + 6 * (size_t)(P.n_fp + q0), bytes_vec, bar);      // b_p (computeScale)
    }
  }
  {  // contribution slots start at zero
    const int coff = T.contrib_off[cta], cn = T.contrib_doubles[cta];
    for (int i = tid; i < cn; i += kTreeThreads) pool[coff + i] = 0.0;
  }
  // this thread's pose for the epilogue (one column per thread; its latency hides behind the solve)
  const int cur = ctl->cur;
  const double lambda = ctl->lambda;
  int kv0 = 0;
  double T0[7];
#pragma unroll
  for (int i = 0; i < 7; ++i) T0[i] = 0.0;
  if (tid < ncols) {
    kv0 = P.pose_of_q[q0 + tid];
#pragma unroll
    for (int i = 0; i < 7; ++i) T0[i] = P.pose[cur][7 * kv0 + i];
  }
  ts_mbar_wait(bar, 0);
  __syncthreads();
  TS_TRACE_G(1);
#ifdef SSBA_SOLVER_TRACE
  int trc = 16;
#endif

  const int *steps = prog + prog[kTH_OffSteps];
  const int nsa = prog[kTH_StepsA], nsb = prog[kTH_StepsB];

  // One instance of the step code for the subtree part and the top part (the kernel is latency bound and one
  // warp often runs alone: instruction fetch matters, the body has to stay small): steps [0, ns_all), with the
  // hand-over of the contributions when the subtree steps are done.
  const int ns_all = nsa + (cta == 0 ? nsb : 0);
  const int ci = 5 * warp + g;

  // ---- numeric factorisation + forward substitution.  Every descriptor a step needs right after a barrier (its
  // table entry, this lane group's diagonal item and first pair word, this warp's first panel item) is fetched one
  // step ahead, so that only loads of numeric operands follow the barriers.
  {
    int4 da = *reinterpret_cast<const int4 *>(steps), db = *reinterpret_cast<const int4 *>(steps + 4);
    int2 dit = make_int2(0, 0);   // this lane group's diagonal item of the running step
    unsigned dpw = 0;             // ... its first pair word
    int2 dpre = make_int2(0, 0);  // ... and the critical panel items of the previous step it scales itself
    int p_n = 0, p_off = 0;       // the previous step's other panel items (rounds, offset)
    if (g < 5 && ci < (da.x & 0xffff)) {
      dit = *reinterpret_cast<const int2 *>(prog + da.y + kTreeItemWords * ci);
      if ((unsigned)dit.x >> 20) dpw = (unsigned)prog[dit.y];
    }
#pragma unroll 1
    for (int s = 0;; ++s) {
      if (s == nsa) {
        TS_TRACE_G(2);
        if (kCluster) {
          // ---- contributions of the subtrees to the top part: up through global memory (L2), added by CTA 0 in
          // CTA order (round r = every destination's r-th contribution: the destinations of a round are distinct)
          if (cta != 0) {
            const double2 *src = reinterpret_cast<const double2 *>(pool + T.contrib_off[cta]);
            double2 *dst = reinterpret_cast<double2 *>(T.xchg + T.xchg_off[cta]);
            const int n2 = T.contrib_doubles[cta] / 2;
            for (int i = tid; i < n2; i += kTreeThreads) dst[i] = src[i];
          }
          ts_cluster_barrier();
          TS_TRACE_G(3);
          if (cta == 0) {
            const int n_rounds = prog[kTH_AddRounds];
            const int *tab = prog + prog[kTH_OffAddRounds];
            constexpr int kAddGroups = kTreeThreads / 18;  // groups of 18 lanes: one 16-byte piece of a block each
            const int grp = tid / 18, sub = tid - 18 * grp;
            const double2 *x2 = reinterpret_cast<const double2 *>(T.xchg);
            for (int rd = 0; rd < n_rounds; ++rd) {
              const int nops = tab[2 * rd];
              const int *ops = prog + tab[2 * rd + 1];
              for (int base = 0; base < nops; base += kAddGroups * 4) {
                double2 v[4];
                int dst[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  const int op = base + kAddGroups * u + grp;
                  dst[u] = -1;
                  if (grp < kAddGroups && op < nops) {
                    const unsigned wd = (unsigned)ops[op];
                    if (sub < ((wd >> 31) ? 3 : 18)) {
                      dst[u] = (int)(wd & 0xffffu) + 2 * sub;
                      v[u] = __ldcg(x2 + ((wd >> 16) & 0x7fffu) + sub);
                    }
                  }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  if (dst[u] >= 0) {
                    double2 *d = reinterpret_cast<double2 *>(pool + dst[u]);
                    double2 t = *d;
                    t.x += v[u].x; t.y += v[u].y;
                    *d = t;
                  }
                }
              }
              __syncthreads();
            }
          }
        }
        TS_TRACE_G(4);
      }
      if (s >= ns_all) break;
      const int nc = da.x & 0xffff, n_look = da.z, off_look = da.w;
      // program data of the next step (nothing the steps write)
      int4 na = make_int4(0, 0, 0, 0), nb = na;
      if (s + 1 < ns_all) { na = *reinterpret_cast<const int4 *>(steps + kTS_Words * (s + 1)); nb = *reinterpret_cast<const int4 *>(steps + kTS_Words * (s + 1) + 4); }
      int2 nit = make_int2(0, 0), npre = make_int2(0, 0);  // the next step's diagonal item, first pair word, critical panel list
      unsigned npw = 0;
      if (g < 5 && ci < (na.x & 0xffff)) {
        nit = *reinterpret_cast<const int2 *>(prog + na.y + kTreeItemWords * ci);
        npre = *reinterpret_cast<const int2 *>(prog + nb.w + 2 * ci);
        if ((unsigned)nit.x >> 20) npw = (unsigned)prog[nit.y];
      }
      // Y = X M row by row (the unscaled row goes to the scratch copy the products read), w = z M for a vector item
      auto panel_item = [&](int2 it) {
        const int nrows = (it.x >> 16) & 15;
        if (r >= nrows) return;
        double2 *D = reinterpret_cast<double2 *>(pool + (it.x & 0xffff) + 6 * r);
        const double2 *M2 = reinterpret_cast<const double2 *>(pool + (it.y & 0xffff));
        const double2 v0 = D[0], v1 = D[1], v2 = D[2];
        double y[6];
#pragma unroll
        for (int c = 0; c < 6; ++c) {
          const double2 m0 = M2[3 * c], m1 = M2[3 * c + 1], m2 = M2[3 * c + 2];
          y[c] = v0.x * m0.x + v0.y * m0.y + v1.x * m1.x + v1.y * m1.y + v2.x * m2.x + v2.y * m2.y;
        }
        D[0] = make_double2(y[0], y[1]); D[1] = make_double2(y[2], y[3]); D[2] = make_double2(y[4], y[5]);
        if (nrows == 6) {
          double2 *X = reinterpret_cast<double2 *>(pool + ((unsigned)it.y >> 16) + 6 * r);
          X[0] = v0; X[1] = v1; X[2] = v2;
        }
      };
      // The step, ONE block barrier: the columns of the previous step get their blocks scaled (Y = X M) first - the
      // blocks this step's diagonal blocks wait for by the lane groups that own those diagonal blocks, all the others
      // by the look-ahead warps; then the diagonal groups apply their critical products and invert, while the
      // look-ahead warps meet everybody's scaled blocks at named barrier 1 and apply the products of the previous
      // step's columns to everything else.
      TS_TRACE_W(s, 0);
      const int nd = nc >= 5 * kTreeWarps ? kTreeWarps : (nc + 4) / 5;
      const bool all_diag = nd == kTreeWarps;
      const int w0 = all_diag ? warp : warp - nd, nw = all_diag ? kTreeWarps : kTreeWarps - nd;
      if (warp < nd) {
        const bool act = g < 5 && ci < nc;
        if (act) {
          for (int i = 0; i < dpre.x; ++i) panel_item(*reinterpret_cast<const int2 *>(prog + dpre.y + kTreeItemWords * i));
        }
        __syncwarp();
        TS_TRACE_W(s, 1);
        // (measured, not kept: arriving only after the critical products, so that the look-ahead warps' loads do not
        // queue in front of the diagonal blocks' operands: 56.0 -> 59.9 us per solve - the look-ahead is on the path too;
        // requesting the first critical pair's operands (21 x 16 bytes per lane) before the arrive: 68 us - the kernel
        // sat at its 128-register cap and the 42 extra live doubles spilled; a second code path for warps with a
        // single column, in which the four idle lane groups share the critical panel item and product - one entry of
        // the 6x6 result per lane: 50.6 -> 54.3 us, the step body no longer fits the instruction cache next to it)
        if (!all_diag) asm volatile("bar.arrive 1, %0;" ::"n"(kTreeThreads) : "memory");
      }
      if (w0 >= 0 && g < 5) {
        const int *rounds = prog + p_off;
#pragma unroll 1
        for (int rd = w0; rd < p_n; rd += nw) panel_item(*reinterpret_cast<const int2 *>(rounds + kTreeRoundWords * rd + kTreeItemWords * g));
      }
      if (warp < nd) {
        const bool act = g < 5 && ci < nc;
        double d[6];
        if (act) ts_product_rows(pool, prog, dit.x, dit.y, dpw, r, d);
        __syncwarp();
        TS_TRACE_W(s, 2);
        if (act && r == 0) {
          if (block_inverse6(pool + (dit.x & 0xffff))) { *s_fail = 1; ctl->chol_fail = 1; }
        }
        TS_TRACE_W(s, 3);
      }
      if (w0 >= 0) {
        TS_TRACE_W(s, 1);
        asm volatile("bar.sync 1, %0;" ::"n"(kTreeThreads) : "memory");
        TS_TRACE_W(s, 2);
        if (n_look > 0) {
          const int *rounds = prog + off_look;
#pragma unroll 1
          for (int rd = w0; rd < n_look; rd += nw) {
            if (g < 5) {
              const int2 it = *reinterpret_cast<const int2 *>(rounds + kTreeRoundWords * rd + kTreeItemWords * g);
              if (r < ((it.x >> 16) & 15)) {
                double d[6];
                ts_product_rows(pool, prog, it.x, it.y, (unsigned)prog[it.y], r, d);
              }
            }
          }
        }
      }
      p_n = db.x; p_off = db.y;  // this step's other panel items: the next step's (or the drain's) business
      da = na; db = nb; dit = nit; dpw = npw; dpre = npre;
      if (w0 >= 0) TS_TRACE_W(s, 3);
      __syncthreads();
      TS_TRACE_W(s, 4);
#ifdef SSBA_SOLVER_TRACE
      TS_TRACE_C(trc); ++trc;
#endif
    }
    // the panel items of the last step (right-hand sides of the root columns)
    if (p_n > 0) {
      if (g < 5) {
        const int *rounds = prog + p_off;
        for (int rd = warp; rd < p_n; rd += kTreeWarps) {
          const int2 it = *reinterpret_cast<const int2 *>(rounds + kTreeRoundWords * rd + kTreeItemWords * g);
          const int nrows = (it.x >> 16) & 15;
          if (r >= nrows) continue;
          double2 *D = reinterpret_cast<double2 *>(pool + (it.x & 0xffff) + 6 * r);
          const double2 *M2 = reinterpret_cast<const double2 *>(pool + (it.y & 0xffff));
          const double2 v0 = D[0], v1 = D[1], v2 = D[2];
          double y[6];
#pragma unroll
          for (int c = 0; c < 6; ++c) {
            const double2 m0 = M2[3 * c], m1 = M2[3 * c + 1], m2 = M2[3 * c + 2];
            y[c] = v0.x * m0.x + v0.y * m0.y + v1.x * m1.x + v1.y * m1.y + v2.x * m2.x + v2.y * m2.y;
          }
          D[0] = make_double2(y[0], y[1]); D[1] = make_double2(y[2], y[3]); D[2] = make_double2(y[4], y[5]);
        }
      }
      __syncthreads();
    }
  }
  TS_TRACE_G(5);

  // ---- backward substitution, right-looking, last step first: the columns i of a step are final when its rounds
  // run; every own column j they reach gets w_j -= Y_ij^T x_i from one lane group (an item per destination, the
  // destinations of the step before first), so that between two steps only one 6x6 product is on the chain.  The
  // top solution goes down to the other CTAs when the top steps are done; their rounds with the top columns as
  // sources come first.
  {
    // (measured, not kept: fetching a step's descriptors - with or without the first item's column of Y_ij - ahead of
    // the barrier that makes its sources final: 49.7 -> 51.6 us per solve at cfg3)
    auto bwd_rounds = [&](int n_rounds, int off) {
      const int *items = prog + off;
#pragma unroll 1
      for (int rd = warp; rd < n_rounds; rd += kTreeWarps) {
        if (g < 5) {
          const int2 it = *reinterpret_cast<const int2 *>(items + kTreeRoundWords * rd + kTreeItemWords * g);
          const int n = (int)((unsigned)it.x >> 16);
          if (n > 0) {
            double acc = 0.0;
            unsigned pw = (unsigned)prog[it.y];
            for (int p = 0; p < n; ++p) {
              const double *B = pool + (pw & 0xffffu);
              const double2 *x2 = reinterpret_cast<const double2 *>(pool + (pw >> 16));
              if (p + 1 < n) pw = (unsigned)prog[it.y + p + 1];
              const double2 x0 = x2[0], x1 = x2[1], xx2 = x2[2];
              acc += B[r] * x0.x + B[6 + r] * x0.y + B[12 + r] * x1.x + B[18 + r] * x1.y + B[24 + r] * xx2.x + B[30 + r] * xx2.y;
            }
            pool[(it.x & 0xffff) + r] -= acc;
          }
        }
      }
    };
#pragma unroll 1
    for (int s = ns_all - 1;; --s) {
      if (s == nsa - 1) {
        TS_TRACE_G(6);
        if (kCluster) {
          if (cta == 0) {
            const int t0 = prog[kTH_TopCol0];
            for (int i = tid + 6 * t0; i < 6 * ncols; i += kTreeThreads) P.xp[6 * (size_t)q0 + i] = pool[V0 + i];
          }
          ts_cluster_barrier();
          if (cta != 0) {
            const int nx = prog[kTH_NXload];
            const int *xl = prog + prog[kTH_OffXload];
            for (int i = tid; i < 6 * nx; i += kTreeThreads) {
              const unsigned wd = (unsigned)xl[i / 6];
              const int m = i - 6 * (i / 6);
              pool[(wd & 0xffffu) + m] = __ldcg(P.xp + 6 * (size_t)(wd >> 16) + m);
            }
            __syncthreads();
            bwd_rounds(prog[kTH_NTopBwd], prog[kTH_OffTopBwd]);
          }
          __syncthreads();
        }
        TS_TRACE_G(7);
      }
      if (s < 1) break;  // the columns of step 0 reach nothing
      bwd_rounds((int)((unsigned)steps[kTS_Words * s + kTS_Cols] >> 16), steps[kTS_Words * s + kTS_OffBwd]);
      __syncthreads();
#ifdef SSBA_SOLVER_TRACE
      TS_TRACE_C(trc); ++trc;
#endif
    }
  }
  TS_TRACE_G(8);

  // ---- epilogue: x_p, the pose part of computeScale and of update(): T <- exp(x) T into the trial buffer
  bool fail;
  if (kCluster) fail = __ldcg(reinterpret_cast<const int *>(&ctl->chol_fail)) != 0;  // raised before the barriers above
  else fail = *s_fail != 0;
  double sc = 0.0;
  for (int t = tid; t < ncols; t += kTreeThreads) {
    int kv = kv0;
    double Tc[7], d[6], out[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) Tc[i] = T0[i];
    if (t != tid) {
      kv = P.pose_of_q[q0 + t];
#pragma unroll
      for (int i = 0; i < 7; ++i) Tc[i] = P.pose[cur][7 * kv + i];
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      d[i] = fail ? 0.0 : pool[V0 + 6 * t + i];
      P.xp[6 * (size_t)(q0 + t) + i] = d[i];
      sc += d[i] * (lambda * d[i] + pool[BP0 + 6 * t + i]);
    }
    pose_oplus(Tc, d, out);
#pragma unroll
    for (int i = 0; i < 7; ++i) P.pose[cur ^ 1][7 * kv + i] = fail ? Tc[i] : out[i];
  }
  sc = ts_block_sum<kTreeThreads>(sc, red);
  if (tid == 0) ctl->scale_pose_part[cta] = sc;
  TS_TRACE_G(9);
}

#ifdef SSBA_SOLVER_TRACE
}  // namespace
extern "C" int ssba_debug_tree_trace(long long *out) {
  cudaDeviceSynchronize();
  return (int)cudaMemcpyFromSymbol(out, g_tree_trace, sizeof(long long) * kTreeMaxCluster * 256);
}
namespace {
#endif

}  // namespace

void fill_tree_dev(const TreeProgram &tp, const int32_t *d_prog, double *d_xchg, TreeDev &out) {
  out = TreeDev{};
  if (!tp.ok) return;
  out.C = tp.C;
  out.smem_bytes = (unsigned)tp.smem_bytes;
  out.prog = d_prog;
  out.xchg = d_xchg;
  for (int c = 0; c <= kTreeMaxCluster; ++c) out.prog_ptr[c] = tp.prog_ptr[c];
  for (int c = 0; c < kTreeMaxCluster; ++c) {
    out.pool_doubles[c] = tp.pool_doubles[c]; out.b0[c] = tp.b0[c]; out.n_own_blocks[c] = tp.n_own_blocks[c];
    out.q0[c] = tp.q0[c]; out.n_own_cols[c] = tp.n_own_cols[c]; out.contrib_off[c] = tp.contrib_off[c];
    out.contrib_doubles[c] = tp.contrib_doubles[c]; out.xchg_off[c] = tp.xchg_off[c];
  }
}

namespace {
// function attributes are per device; one mutex keeps the set-up and the first launches of other threads apart
void tree_setup_device() {
  static std::atomic<unsigned long long> done{0};
  static std::atomic_flag lock = ATOMIC_FLAG_INIT;
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned long long bit = 1ull << (dev & 63);
  if (done.load(std::memory_order_acquire) & bit) return;
  while (lock.test_and_set(std::memory_order_acquire)) { }
  if (!(done.load(std::memory_order_acquire) & bit)) {
    cudaFuncSetAttribute(k_tree_solve<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTreeMaxSmem);
    cudaFuncSetAttribute(k_tree_solve<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTreeMaxSmem);
    cudaFuncSetAttribute(k_tree_solve<true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaGetLastError();
    done.fetch_or(bit, std::memory_order_release);
  }
  lock.clear(std::memory_order_release);
}
}  // namespace

// The largest cluster (16, 8, 4, 2 or 1 CTAs with the full dynamic shared memory) the device can co-schedule
int max_tree_cluster() {
  tree_setup_device();
  for (int c = kTreeMaxCluster; c >= 2; c /= 2) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(c); cfg.blockDim = dim3(kTreeThreads); cfg.dynamicSmemBytes = kTreeMaxSmem;
    cudaLaunchAttribute attr{};
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = c; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, k_tree_solve<true>, &cfg) == cudaSuccess && n >= 1) return c;
    cudaGetLastError();
  }
  return 1;
}

void launch_tree_solve(const DeviceProblem &P, cudaStream_t st) {
  tree_setup_device();
  const int c = P.tree.C;
  if (c == 1) { launch_maybe_pdl(k_tree_solve<false>, dim3(1), dim3(kTreeThreads), P.tree.smem_bytes, st, P.pdl != 0, P); return; }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(c); cfg.blockDim = dim3(kTreeThreads); cfg.dynamicSmemBytes = P.tree.smem_bytes; cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = c; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = P.pdl ? 2 : 1;
  cudaLaunchKernelEx(&cfg, k_tree_solve<true>, P);
}

}  // namespace ssba
