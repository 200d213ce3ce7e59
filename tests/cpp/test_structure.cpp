// Host-only check of the structure builder (ssvio_b200/csrc/ssba_structure.cpp): runs the
// block-sparse left-looking Cholesky + level-ordered substitutions exactly as the CUDA kernel
// walks them (upd lists, row lists, levels), but on the CPU, over a random SPD matrix with the
// Schur pattern of a synthetic graph, and compares with a dense solve.  Also checks the pair /
// chunk bookkeeping invariants.  Built and run by tests/test_structure.py (no GPU needed).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "ssba_structure.hpp"
#include "ssba_block_inverse.cuh"

using namespace ssba;

static int fails = 0;
#define CHECK(c, ...) do { if (!(c)) { std::printf("FAIL %s:%d: ", __FILE__, __LINE__); std::printf(__VA_ARGS__); std::printf("\n"); ++fails; } } while (0)

static HostGraph make_graph(int nk, int np, int w, unsigned seed, bool fix0, int n_fixed_pts, bool loop) {
  HostGraph g;
  std::mt19937 rng(seed);
  g.have_cams = true; g.cams.n = 2;
  g.n_poses = nk; g.n_points = np;
  g.poses.assign(7 * nk, 0.0); g.points.assign(3 * np, 0.0);
  g.pose_fixed.assign(nk, 0); g.point_fixed.assign(np, 0);
  if (fix0) g.pose_fixed[0] = 1;
  for (int i = 0; i < n_fixed_pts && i < np; ++i) g.point_fixed[(i * 7919) % np] = 1;
  for (int j = 0; j < np; ++j) {
    int k0 = rng() % (nk - w + 1);
    for (int k = 0; k < w; ++k)
      for (int c = 0; c < 2; ++c) {
        int pose = k0 + k;
        if (loop && (j % 17 == 0) && k == w - 1) pose = (k0 + nk / 2) % nk;  // long-range covisibility
        g.e_pose.push_back(pose); g.e_point.push_back(j); g.e_cam.push_back(c);
        g.e_uv.push_back(0); g.e_uv.push_back(0);
      }
  }
  g.n_edges = (int)g.e_pose.size();
  // shuffle edge order to exercise the sort
  std::vector<int> perm(g.n_edges);
  for (int i = 0; i < g.n_edges; ++i) perm[i] = i;
  std::shuffle(perm.begin(), perm.end(), rng);
  HostGraph h = g;
  for (int i = 0; i < g.n_edges; ++i) {
    h.e_pose[i] = g.e_pose[perm[i]]; h.e_point[i] = g.e_point[perm[i]]; h.e_cam[i] = g.e_cam[perm[i]];
  }
  return h;
}

// ---- CPU interpreter of the PACKED device program (Structure::prog), i.e. of what k_reduced_solve
// really walks: per level the rounds of every CTA of the cluster (DIAG / SUB / VEC / look-ahead ACC,
// split REDUCE rounds), block references that are shared-memory slots of the executing CTA or global
// blocks, the reader masks of the DSMEM broadcast, the slot plan (a slot read before it was written, or
// after it was handed to another block, gives a wrong or NaN result), and the backward lists.
// Mirrors LevelSeg in ssba_kernels.cu.  A0 = initial block values (Schur complement), b = right-hand side.
#include "ssba_solver_layout.hpp"
#include <limits>
static std::vector<double> interpret_program(const Structure &s, std::vector<double> L, const std::vector<double> &b) {
  const int n = s.n_fp, C = s.solve_cluster, NSEG = s.n_segments;
  const SolverSmemLayout lay = solver_smem_layout(n, s.prog_max_seg);
  const double nan = std::numeric_limits<double>::quiet_NaN();
  std::vector<std::vector<double>> slots(C, std::vector<double>(36 * (size_t)std::max(lay.n_slots, 1), nan));
  std::vector<double> y(6 * (size_t)n, nan);
  struct Seg { int n_cols, n_rounds, n_pairs, n_brows, n_pf, n_bpf; const int *cta_rptr, *col_j, *col_b0, *col_bptr, *brow, *round_type, *gt_dst, *gt_slot, *gt_pos, *gt_p0, *gt_p1, *gt_mask, *pa, *pb; };
  auto parse = [&](int sg) {
    const int *seg = s.prog.data() + s.prog_ptr[sg];
    Seg S; S.n_cols = seg[0]; S.n_rounds = seg[1]; S.n_pairs = seg[2]; S.n_brows = seg[3]; S.n_pf = seg[4]; S.n_bpf = seg[5];
    const int *p = seg + 8;
    S.cta_rptr = p; p += kSolveMaxCluster + 1;
    S.col_j = p; p += S.n_cols; S.col_b0 = p; p += S.n_cols; S.col_bptr = p; p += S.n_cols + 1; S.brow = p; p += S.n_brows;
    S.round_type = p; p += S.n_rounds;
    S.gt_dst = p; p += 5 * S.n_rounds; S.gt_slot = p; p += 5 * S.n_rounds; S.gt_pos = p; p += 5 * S.n_rounds;
    S.gt_p0 = p; p += 5 * S.n_rounds; S.gt_p1 = p; p += 5 * S.n_rounds; S.gt_mask = p; p += 5 * S.n_rounds;
    S.pa = p; p += S.n_pairs; S.pb = p; p += S.n_pairs;
    CHECK(p + 2 * S.n_pf + S.n_bpf <= s.prog.data() + s.prog_ptr[sg + 1], "segment %d overruns", sg);
    return S;
  };
  for (int sg = 1; sg < NSEG; ++sg) {
    const Seg S = parse(sg);
    CHECK(S.cta_rptr[0] == 0 && S.cta_rptr[C] == S.n_rounds, "round ranges of level %d", sg - 1);
    std::vector<std::vector<double>> linv(C, std::vector<double>(36 * (size_t)std::max(S.n_cols, 1), nan));
    auto blk = [&](int cta, int ref) -> const double * { return ref >= 0 ? &slots[cta][36 * (size_t)ref] : &L[36 * (size_t)(-1 - ref)]; };
    auto products = [&](int cta, int p0, int p1, double *acc) {  // acc += sum of A B^T
      for (int p = p0; p < p1; ++p) {
        const double *A = blk(cta, S.pa[p]), *B = blk(cta, S.pb[p]);
        for (int r = 0; r < 6; ++r) for (int c = 0; c < 6; ++c) { double t = 0; for (int k = 0; k < 6; ++k) t += A[6 * r + k] * B[6 * c + k]; acc[6 * r + c] += t; }
      }
    };
    // the device runs the DIAG rounds of a CTA before its other rounds; the host order has them first
    for (int pass = 0; pass < 2; ++pass)
      for (int cta = 0; cta < C; ++cta)
        for (int rd = S.cta_rptr[cta]; rd < S.cta_rptr[cta + 1]; ++rd) {
          const int kind = S.round_type[rd] & 3;
          const bool reduce = (S.round_type[rd] & 4) != 0;
          if ((kind == 0) != (pass == 0)) continue;
          for (int g = 0; g < (reduce ? 1 : 5); ++g) {
            const int gt = 5 * rd + g, dst = S.gt_dst[gt], pos = S.gt_pos[gt];
            if (dst < 0) continue;
            if (kind == 2) {  // VEC: y_j = Linv (b_j - sum L(j,k) y_k)
              const int j = S.col_j[pos];
              double sv[6];
              for (int r = 0; r < 6; ++r) sv[r] = b[6 * j + r];
              for (int gg = 0; gg < (reduce ? 5 : 1); ++gg)
                for (int p = S.gt_p0[5 * rd + (reduce ? gg : g)]; p < S.gt_p1[5 * rd + (reduce ? gg : g)]; ++p) {
                  const double *B = blk(cta, S.pa[p]); const double *yk = &y[6 * (size_t)S.pb[p]];
                  for (int r = 0; r < 6; ++r) for (int c = 0; c < 6; ++c) sv[r] -= B[6 * r + c] * yk[c];
                }
              const double *Li = &linv[cta][36 * (size_t)pos];
              for (int r = 0; r < 6; ++r) { double t = 0; for (int c = 0; c <= r; ++c) t += Li[6 * r + c] * sv[c]; y[6 * j + r] = t; }
              continue;
            }
            double acc[36] = {0};
            for (int gg = 0; gg < (reduce ? 5 : 1); ++gg) products(cta, S.gt_p0[5 * rd + (reduce ? gg : g)], S.gt_p1[5 * rd + (reduce ? gg : g)], acc);
            double v[36];
            for (int i = 0; i < 36; ++i) v[i] = L[36 * (size_t)dst + i] - acc[i];
            if (kind == 3) { for (int i = 0; i < 36; ++i) L[36 * (size_t)dst + i] = v[i]; continue; }  // ACC
            if (kind == 0) {  // DIAG: Cholesky of the lower triangle, explicit inverse
              double a[36] = {0}, X[36] = {0};
              for (int i = 0; i < 6; ++i) for (int k = 0; k <= i; ++k) a[6 * i + k] = v[6 * i + k];
              for (int j = 0; j < 6; ++j) {
                CHECK(a[7 * j] > 0, "pivot of block %d", dst);
                const double d = std::sqrt(a[7 * j]); a[7 * j] = d;
                for (int i = j + 1; i < 6; ++i) a[6 * i + j] /= d;
                for (int i = j + 1; i < 6; ++i) for (int k = j + 1; k <= i; ++k) a[6 * i + k] -= a[6 * i + j] * a[6 * k + j];
              }
              for (int c = 0; c < 6; ++c) {
                X[7 * c] = 1.0 / a[7 * c];
                for (int i = c + 1; i < 6; ++i) { double t = 0; for (int k = c; k < i; ++k) t += a[6 * i + k] * X[6 * k + c]; X[6 * i + c] = -t / a[7 * i]; }
              }
              for (int i = 0; i < 36; ++i) { L[36 * (size_t)dst + i] = X[i]; linv[cta][36 * (size_t)pos + i] = X[i]; }
            } else {  // SUB: X = v Linv^T, stored in global memory and in the caches of the reader CTAs
              const double *Li = &linv[cta][36 * (size_t)pos];
              double X[36];
              for (int r = 0; r < 6; ++r) for (int c = 0; c < 6; ++c) { double t = 0; for (int k = 0; k <= c; ++k) t += v[6 * r + k] * Li[6 * c + k]; X[6 * r + c] = t; }
              for (int i = 0; i < 36; ++i) L[36 * (size_t)dst + i] = X[i];
              const int slot = S.gt_slot[gt];
              if (slot >= 0) {
                CHECK(slot < lay.n_slots, "slot out of range");
                for (int cc = 0; cc < C; ++cc) if (C == 1 || ((S.gt_mask[gt] >> cc) & 1)) for (int i = 0; i < 36; ++i) slots[cc][36 * (size_t)slot + i] = X[i];
              }
            }
          }
        }
  }
  // backward: x_j = Linv_jj^T (y_j - sum_i L(i,j)^T x_i)
  std::vector<double> x = y;
  for (int sg = NSEG - 1; sg >= 1; --sg) {
    const Seg S = parse(sg);
    for (int t = 0; t < S.n_cols; ++t) {
      const int j = S.col_j[t], b0 = S.col_b0[t], nb = S.col_bptr[t + 1] - S.col_bptr[t];
      const int *rows = S.brow + S.col_bptr[t];
      double sv[6];
      for (int r = 0; r < 6; ++r) sv[r] = x[6 * j + r];
      for (int k = 0; k < nb; ++k) { const double *B = &L[36 * (size_t)(b0 + 1 + k)]; for (int r = 0; r < 6; ++r) for (int c = 0; c < 6; ++c) sv[r] -= B[6 * c + r] * x[6 * rows[k] + c]; }
      const double *Li = &L[36 * (size_t)b0];
      for (int r = 0; r < 6; ++r) { double t2 = 0; for (int c = r; c < 6; ++c) t2 += Li[6 * c + r] * sv[c]; x[6 * j + r] = t2; }
    }
  }
  return x;
}

// ---- CPU interpreter of the subtree-per-CTA program (Structure::tree, ssba_tree_program.cpp), i.e. of what
// k_tree_solve walks: per CTA its shared-memory pool (own factor blocks, vectors, contribution slots), per step
// the diagonal items (critical products + inverse of the 6x6 block), the look-ahead product rounds that run beside
// them, the panel rounds (Y = X M in place, the unscaled X into the one-step scratch copy); the contribution hand-off to CTA 0 (add rounds), the top part, the backward passes and the
// hand-off of the top solution.  Every pool double remembers who wrote / read it in the running barrier
// interval: two different work items touching the same double in one interval (one of them writing) is a race
// on the device and fails here.
struct TreeCta {
  std::vector<double> pool;
  std::vector<int> wr_t, wr_i, rd_t, rd_i;
  // the diagonal items of a step are still running while the look-ahead items of the same step run: what they
  // touch (dg_mode 1: record under dg_stamp) must not be touched by a look-ahead item (dg_mode 2: check)
  std::vector<int> dg_w, dg_r;
  int dg_mode = 0, dg_stamp = 0;
  const int32_t *w;
  int T = 0, item = 0;
  double rd2(int off) { if (off & 1) CHECK(false, "tree program: odd offset %d under a 16-byte access", off); return 0.0; }
  double rd(int off) {
    if (dg_mode == 1) dg_r[off] = dg_stamp;
    if (dg_mode == 2 && dg_w[off] == dg_stamp) CHECK(false, "tree program: a look-ahead item reads a double a diagonal item of the same step writes (off %d)", off);
    if (wr_t[off] == T && wr_i[off] != item) { CHECK(false, "tree program: read of a double written by another item in the same interval (off %d)", off); }
    if (rd_t[off] == T && rd_i[off] != item) rd_i[off] = -2; else { rd_t[off] = T; rd_i[off] = item; }
    return pool[off];
  }
  void wrt(int off, double v) {
    if (dg_mode == 1) dg_w[off] = dg_stamp;
    if (dg_mode == 2 && (dg_w[off] == dg_stamp || dg_r[off] == dg_stamp)) CHECK(false, "tree program: a look-ahead item writes a double a diagonal item of the same step touches (off %d)", off);
    if (rd_t[off] == T && rd_i[off] != item) { CHECK(false, "tree program: write of a double read by another item in the same interval (off %d)", off); }
    if (wr_t[off] == T && wr_i[off] != item) { CHECK(false, "tree program: two items write the same double in one interval (off %d)", off); }
    wr_t[off] = T; wr_i[off] = item;
    pool[off] = v;
  }
};
static void tree_products(TreeCta &c, int dest, int nrows, int p0, int p1) {
  for (int r = 0; r < nrows; ++r) {
    double acc[6] = {0, 0, 0, 0, 0, 0};
    for (int p = p0; p < p1; ++p) {
      const unsigned wd = (unsigned)c.w[p];
      const int a = (int)(wd & 0xffff) + 6 * r, b = (int)(wd >> 16);
      c.rd2(a); c.rd2(b); c.rd2(dest);
      for (int m = 0; m < 6; ++m) { double t = 0; for (int k = 0; k < 6; ++k) t += c.rd(a + k) * c.rd(b + 6 * m + k); acc[m] += t; }
    }
    if (p1 > p0) for (int m = 0; m < 6; ++m) c.wrt(dest + 6 * r + m, c.rd(dest + 6 * r + m) - acc[m]);
  }
}
static void tree_panel_item(TreeCta &c, const int32_t *it) {
  const int dest = it[0] & 0xffff, nrows = (it[0] >> 16) & 15, dg = it[1] & 0xffff, xc = (int)((unsigned)it[1] >> 16);
  CHECK(dg % 2 == 0 && xc % 2 == 0 && dest % 2 == 0, "panel item misaligned");
  for (int r = 0; r < nrows; ++r) {
    double x[6], y[6];
    for (int m = 0; m < 6; ++m) x[m] = c.rd(dest + 6 * r + m);
    for (int m = 0; m < 6; ++m) { double v = 0; for (int k = 0; k < 6; ++k) v += x[k] * c.rd(dg + 6 * m + k); y[m] = v; }
    for (int m = 0; m < 6; ++m) c.wrt(dest + 6 * r + m, y[m]);       // Y = X M in place
    if (nrows == 6) for (int m = 0; m < 6; ++m) c.wrt(xc + 6 * r + m, x[m]);  // the unscaled row for the next step's products
  }
}
// One step = two intervals on the device.  Interval A: every diagonal column as ONE work item (the panel items of
// the previous step that feed it, its critical products, the inverse - one lane group does all three, in order)
// beside the other panel items of the previous step (the look-ahead warps).  Named barrier.  Interval B: the
// look-ahead products - the diagonal groups may still be busy with their item, which is why that item must not
// touch anything an interval-B item touches (same T for both is too strict only for reads of the inverse, which
// nobody does before the next step).
static void tree_forward_steps(TreeCta &c, int s0, int s1, bool *fail) {
  const int32_t *w = c.w;
  const int32_t *prev = nullptr;  // the previous step's table entry
  for (int s = s0; s <= s1; ++s) {
    const int32_t *st = s < s1 ? w + w[kTH_OffSteps] + kTS_Words * s : nullptr;
    if (st) CHECK(st[kTS_OffDiag] % 2 == 0 && st[kTS_OffLook] % 2 == 0 && st[kTS_OffPanel] % 2 == 0 && st[kTS_OffBwd] % 2 == 0 && st[kTS_OffPre] % 2 == 0, "tree program: item arrays misaligned");
    ++c.T;  // interval A (after the last step: the drain of its panel items)
    ++c.dg_stamp;
    if (st)
      for (int t = 0; t < (st[kTS_Cols] & 0xffff); ++t) {
        ++c.item;
        const int32_t *pre = w + st[kTS_OffPre] + 2 * t;
        for (int i = 0; i < pre[0]; ++i) tree_panel_item(c, w + pre[1] + kTreeItemWords * i);
        c.dg_mode = 1;  // from here on (after the group's arrival at the named barrier) it runs beside the look-ahead items
        const int32_t *it = w + st[kTS_OffDiag] + kTreeItemWords * t;
        const int d = it[0] & 0xffff;
        CHECK(((it[0] >> 16) & 15) == 6, "diagonal item rows");
        tree_products(c, d, 6, it[1], it[1] + (int)((unsigned)it[0] >> 20));
        // M = D^-1 from the LOWER triangle, written back as the full symmetric block; D must be positive definite
        double a[36], Ls[36] = {0}, Li[36] = {0}, M[36];
        for (int i = 0; i < 6; ++i) for (int k = 0; k <= i; ++k) a[6 * i + k] = c.rd(d + 6 * i + k);
        for (int j = 0; j < 6; ++j) {  // reference route: Cholesky, triangular inverse, M = L^-T L^-1
          double dj = a[7 * j];
          for (int k = 0; k < j; ++k) dj -= Ls[6 * j + k] * Ls[6 * j + k];
          if (!(dj > 0)) { *fail = true; dj = 1.0; }
          Ls[7 * j] = std::sqrt(dj);
          for (int i = j + 1; i < 6; ++i) { double v = a[6 * i + j]; for (int k = 0; k < j; ++k) v -= Ls[6 * i + k] * Ls[6 * j + k]; Ls[6 * i + j] = v / Ls[7 * j]; }
        }
        for (int j = 0; j < 6; ++j) {
          Li[7 * j] = 1.0 / Ls[7 * j];
          for (int i = j + 1; i < 6; ++i) { double v = 0; for (int k = j; k < i; ++k) v -= Ls[6 * i + k] * Li[6 * k + j]; Li[6 * i + j] = v / Ls[7 * i]; }
        }
        for (int i = 0; i < 6; ++i) for (int k = 0; k < 6; ++k) { double v = 0; for (int m = 0; m < 6; ++m) v += Li[6 * m + i] * Li[6 * m + k]; M[6 * i + k] = v; }
        {  // the product's closed-form inverse (ssba_block_inverse.cuh) is what the device runs: use it, and compare
          double Dm[36];
          for (int i = 0; i < 6; ++i) for (int k = 0; k < 6; ++k) Dm[6 * i + k] = k <= i ? a[6 * i + k] : std::numeric_limits<double>::quiet_NaN();  // upper triangle must not be read
          const bool bad = block_inverse6(Dm);
          if (bad) *fail = true;
          double scale = 0; for (int i = 0; i < 36; ++i) scale = std::max(scale, std::fabs(M[i]));
          for (int i = 0; i < 36; ++i) CHECK(std::fabs(Dm[i] - M[i]) <= 1e-11 * scale, "closed-form inverse vs Cholesky route: %g vs %g", Dm[i], M[i]);
          for (int i = 0; i < 36; ++i) M[i] = Dm[i];
        }
        for (int i = 0; i < 36; ++i) c.wrt(d + i, M[i]);
        c.dg_mode = 0;
      }
    if (prev)
      for (int i = 0; i < 5 * prev[kTS_NPanel]; ++i) {
        ++c.item;
        tree_panel_item(c, w + prev[kTS_OffPanel] + kTreeItemWords * i);
      }
    if (!st) break;
    // interval B (after named barrier 1): look-ahead rounds; the diagonal items of interval A may still be running
    ++c.T;
    c.dg_mode = 2;
    for (int i = 0; i < 5 * st[kTS_NLook]; ++i) {
      ++c.item;
      const int32_t *it = w + st[kTS_OffLook] + kTreeItemWords * i;
      const int nrows = (it[0] >> 16) & 15;
      if (nrows > 0) tree_products(c, it[0] & 0xffff, nrows, it[1], it[1] + (int)((unsigned)it[0] >> 20));
    }
    c.dg_mode = 0;
    prev = st;
  }
}
// the backward rounds of one source step (or of the top columns for a CTA != 0): every item is one work item of the
// interval; it reads final x_i, blocks Y_ij and rewrites its own destination w_j
static void tree_backward_rounds(TreeCta &c, int n_rounds, int off) {
  const int32_t *w = c.w;
  CHECK(off % 2 == 0, "tree program: backward items misaligned");
  ++c.T;
  for (int i = 0; i < 5 * n_rounds; ++i) {
    ++c.item;
    const int32_t *it = w + off + kTreeItemWords * i;
    const int dest = it[0] & 0xffff, n = (int)((unsigned)it[0] >> 16);
    if (n == 0) continue;
    double acc[6] = {0, 0, 0, 0, 0, 0};
    for (int p = 0; p < n; ++p) {
      const unsigned pw = (unsigned)w[it[1] + p];
      const int B = (int)(pw & 0xffff), xo = (int)(pw >> 16);
      c.rd2(xo);
      for (int m = 0; m < 6; ++m) { double t = 0; for (int nn = 0; nn < 6; ++nn) t += c.rd(B + 6 * nn + m) * c.rd(xo + nn); acc[m] += t; }
    }
    for (int m = 0; m < 6; ++m) c.wrt(dest + m, c.rd(dest + m) - acc[m]);  // w_j -= sum_i Y_ij^T x_i
  }
}
static void tree_backward_steps(TreeCta &c, int s0, int s1) {
  const int32_t *w = c.w;
  for (int s = s1 - 1; s >= std::max(s0, 1); --s) {
    const int32_t *st = w + w[kTH_OffSteps] + kTS_Words * s;
    tree_backward_rounds(c, (int)((unsigned)st[kTS_Cols] >> 16), st[kTS_OffBwd]);
  }
}
static std::vector<double> interpret_tree_program(const Structure &s, const std::vector<double> &L0, const std::vector<double> &b) {
  const TreeProgram &tp = s.tree;
  const int n = s.n_fp, C = tp.C;
  const double nan = std::numeric_limits<double>::quiet_NaN();
  std::vector<TreeCta> cta(C);
  std::vector<double> xchg((size_t)std::max(tp.xchg_doubles, 1), nan), x(6 * (size_t)n, nan);
  CHECK(tp.smem_bytes <= kTreeMaxSmem, "tree program: %zu bytes of shared memory", tp.smem_bytes);
  int cols_seen = 0;
  for (int c = 0; c < C; ++c) {
    TreeCta &t = cta[c];
    t.w = tp.words.data() + tp.prog_ptr[c];
    const int np = tp.pool_doubles[c];
    CHECK(8 * (size_t)np + 4 * (size_t)(tp.prog_ptr[c + 1] - tp.prog_ptr[c]) + kTreeMiscBytes <= tp.smem_bytes, "tree program: CTA %d exceeds the launch size", c);
    t.pool.assign(np, nan); t.wr_t.assign(np, -1); t.wr_i.assign(np, -1); t.rd_t.assign(np, -1); t.rd_i.assign(np, -1);
    t.dg_w.assign(np, -1); t.dg_r.assign(np, -1);
    for (int i = 0; i < 36 * tp.n_own_blocks[c]; ++i) t.pool[i] = L0[36 * (size_t)tp.b0[c] + i];
    for (int i = 0; i < 6 * tp.n_own_cols[c]; ++i) t.pool[36 * tp.n_own_blocks[c] + i] = b[6 * (size_t)tp.q0[c] + i];
    for (int i = 0; i < tp.contrib_doubles[c]; ++i) t.pool[tp.contrib_off[c] + i] = 0.0;
    cols_seen += tp.n_own_cols[c];
  }
  CHECK(cols_seen == n, "tree program: CTAs own %d of %d columns", cols_seen, n);
  bool fail = false;
  for (int c = 0; c < C; ++c) tree_forward_steps(cta[c], 0, cta[c].w[kTH_StepsA], &fail);
  for (int c = 1; c < C; ++c)
    for (int i = 0; i < tp.contrib_doubles[c]; ++i) xchg[tp.xchg_off[c] + i] = cta[c].pool[tp.contrib_off[c] + i];
  {
    TreeCta &t = cta[0];
    const int32_t *w = t.w;
    for (int r = 0; r < w[kTH_AddRounds]; ++r) {
      ++t.T;
      const int nops = w[w[kTH_OffAddRounds] + 2 * r], off = w[w[kTH_OffAddRounds] + 2 * r + 1];
      for (int i = 0; i < nops; ++i) {
        ++t.item;
        const unsigned op = (unsigned)w[off + i];
        const int dest = (int)(op & 0xffff), src = 2 * (int)((op >> 16) & 0x7fff), nd = (op >> 31) ? 6 : 36;
        for (int k = 0; k < nd; ++k) t.wrt(dest + k, t.rd(dest + k) + xchg[src + k]);
      }
    }
    tree_forward_steps(t, w[kTH_StepsA], w[kTH_StepsA] + w[kTH_StepsB], &fail);
    tree_backward_steps(t, w[kTH_StepsA], w[kTH_StepsA] + w[kTH_StepsB]);
  }
  CHECK(!fail, "tree program: pivot <= 0");
  // the top solution goes down through global memory
  {
    const int V0 = 36 * tp.n_own_blocks[0];
    for (int q = tp.q0[0]; q < n; ++q) for (int m = 0; m < 6; ++m) x[6 * (size_t)q + m] = cta[0].pool[V0 + 6 * (q - tp.q0[0]) + m];
  }
  for (int c = 1; c < C; ++c) {
    TreeCta &t = cta[c];
    ++t.T; ++t.item;
    for (int i = 0; i < t.w[kTH_NXload]; ++i) {
      const unsigned wd = (unsigned)t.w[t.w[kTH_OffXload] + i];
      for (int m = 0; m < 6; ++m) t.wrt((int)(wd & 0xffff) + m, x[6 * (size_t)(wd >> 16) + m]);
    }
    tree_backward_rounds(t, t.w[kTH_NTopBwd], t.w[kTH_OffTopBwd]);
  }
  for (int c = 0; c < C; ++c) {
    tree_backward_steps(cta[c], 0, cta[c].w[kTH_StepsA]);
    const int V0 = 36 * tp.n_own_blocks[c];
    for (int q = tp.q0[c]; q < tp.q0[c] + tp.n_own_cols[c]; ++q) for (int m = 0; m < 6; ++m) x[6 * (size_t)q + m] = cta[c].pool[V0 + 6 * (q - tp.q0[c]) + m];
  }
  return x;
}

static void check_case(int nk, int np, int w, unsigned seed, bool fix0, int nfix, bool loop, int world) {
  HostGraph g = make_graph(nk, np, w, seed, fix0, nfix, loop);
  std::vector<Structure> S(world);
  std::string err;
  long long edges_seen = 0, slots_free = 0;
  for (int r = 0; r < world; ++r) {
    bool ok = build_structure(g, r, world, S[r], err);
    CHECK(ok, "build_structure: %s", err.c_str());
    if (!ok) return;
    edges_seen += S[r].n_edges;
    slots_free += S[r].n_fl;
    // pairs: W-pairs first sorted by q, combos point at existing blocks
    const Structure &s = S[r];
    for (int sl = 0; sl < s.n_slots; ++sl) {
      int prev = -1; bool seen_fixed = false; int k = 0;
      for (int a = s.slot_pair_ptr[sl]; a < s.slot_pair_ptr[sl + 1]; ++a) {
        CHECK(s.pair_edge_ptr[a + 1] > s.pair_edge_ptr[a], "empty pair");
        if (s.pair_q[a] < 0) { seen_fixed = true; continue; }
        CHECK(!seen_fixed, "free-pose pair after fixed-pose pair");
        CHECK(s.pair_q[a] > prev, "pairs not sorted by q"); prev = s.pair_q[a]; ++k;
        CHECK(s.q_of_pose[s.pair_vertex[a]] == s.pair_q[a], "pair_q mismatch");
      }
    }
    // Schur units: every free landmark is in exactly one run; the units of a run cover its
    // k(k+1)/2 block pairs once, and their targets are the blocks (row q_b, col q_a)
    {
      std::vector<int> covered(s.n_slots, 0);
      for (int u = 0; u < s.n_units; ++u) {
        const int s0 = s.unit_slot[u], nrun = s.unit_n[u], k = s.unit_k[u], c0 = s.unit_c0[u];
        const int nc = s.unit_combo_ptr[u + 1] - s.unit_combo_ptr[u];
        CHECK(nc == std::min(32, k * (k + 1) / 2 - c0) && nc > 0, "unit %d: %d targets", u, nc);
        if (c0 == 0) for (int i = 0; i < nrun; ++i) ++covered[s0 + i];
        for (int i = 0; i < nrun; ++i) {
          CHECK(s.slot_free[s0 + i], "fixed landmark in a Schur run");
          for (int a = 0; a < k; ++a)
            CHECK(s.pair_q[s.slot_pair_ptr[s0 + i] + a] == s.pair_q[s.slot_pair_ptr[s0] + a], "run with different pose lists");
        }
        int idx = 0;
        for (int a = 0; a < k; ++a)
          for (int b2 = a; b2 < k; ++b2, ++idx) {
            if (idx < c0 || idx >= c0 + nc) continue;
            const int b = s.combo_blk[s.unit_combo_ptr[u] + idx - c0];
            CHECK(s.blk_col[b] == s.pair_q[s.slot_pair_ptr[s0] + a] && s.blk_row[b] == s.pair_q[s.slot_pair_ptr[s0] + b2], "combo block wrong");
          }
      }
      // producer lists of the deterministic accumulation: every combo appears once, under its block, in unit order
      CHECK((int)s.blk_prod_ptr.size() == s.n_blocks + 1 && s.blk_prod.size() == s.combo_blk.size(), "producer lists sized wrong");
      if ((int)s.blk_prod_ptr.size() == s.n_blocks + 1) {
        std::vector<int> seen_combo(s.combo_blk.size(), 0);
        for (int b = 0; b < s.n_blocks; ++b)
          for (int p = s.blk_prod_ptr[b]; p < s.blk_prod_ptr[b + 1]; ++p) {
            const int c = s.blk_prod[p];
            CHECK(s.combo_blk[c] == b, "producer %d listed under block %d but targets %d", c, b, s.combo_blk[c]);
            CHECK(p == s.blk_prod_ptr[b] || s.blk_prod[p - 1] < c, "producers of block %d out of order", b);
            ++seen_combo[c];
          }
        for (int v : seen_combo) CHECK(v == 1, "combo listed %d times", v);
      }
      for (int sl = 0; sl < s.n_slots; ++sl) {
        int k = 0;
        for (int a = s.slot_pair_ptr[sl]; a < s.slot_pair_ptr[sl + 1]; ++a) if (s.pair_q[a] >= 0) ++k;
        CHECK(covered[sl] == ((s.slot_free[sl] && k > 0) ? 1 : 0), "landmark %d in %d Schur runs", sl, covered[sl]);
      }
    }
    for (int e = 0; e < s.n_edges; ++e) CHECK(s.e_orig[e] >= 0 && s.e_orig[e] < g.n_edges, "e_orig range");
    // Hpp partial bookkeeping: every free-pose pair is in exactly one partial of its chunk, and the
    // partial lists of the poses cover all partials
    {
      long long covered = 0, free_pairs = 0;
      for (int a = 0; a < s.n_pairs; ++a) if (s.pair_q[a] >= 0) ++free_pairs;
      std::vector<int> seen(s.n_hpp_parts, 0);
      for (int q = 0; q < s.n_fp; ++q)
        for (int i = s.q_part_ptr[q]; i < s.q_part_ptr[q + 1]; ++i) ++seen[s.q_part[i]];
      for (int v : seen) CHECK(v == 1, "partial owned by %d poses", v);
      for (int c = 0; c < s.n_lchunks; ++c) {
        const int a0 = s.slot_pair_ptr[s.lchunk_slot[c]], a1 = s.slot_pair_ptr[s.lchunk_slot[c + 1]];
        if (a1 - a0 > 128) { CHECK(s.lchunk_slot[c + 1] - s.lchunk_slot[c] == 1, "big chunk with several landmarks"); for (int a = a0; a < a1; ++a) if (s.pair_q[a] >= 0) ++covered; continue; }
        for (int lp = s.lchunk_lp_ptr[c]; lp < s.lchunk_lp_ptr[c + 1]; ++lp) {
          int q = -2;
          for (int i = s.lp_pair_ptr[lp]; i < s.lp_pair_ptr[lp + 1]; ++i) {
            const int a = a0 + s.lp_pair[i];
            CHECK(a < a1 && s.pair_q[a] >= 0, "lp_pair out of chunk");
            if (q == -2) q = s.pair_q[a];
            CHECK(s.pair_q[a] == q, "mixed poses in one partial");
            ++covered;
          }
        }
      }
      CHECK(covered == free_pairs, "partials cover %lld of %lld free-pose pairs", covered, free_pairs);
    }

  }
  CHECK(edges_seen == S[0].n_active_edges_global, "shards cover %lld of %d active edges", edges_seen, S[0].n_active_edges_global);
  CHECK(slots_free == S[0].n_fl_global, "free landmarks %lld vs %d", slots_free, S[0].n_fl_global);

  // numeric factorisation through the structure arrays vs dense
  const Structure &s = S[0];
  const int n = s.n_fp, N = 6 * n;
  if (n == 0) return;
  std::mt19937 rng(seed + 1);
  std::uniform_real_distribution<double> U(-1, 1);
  std::vector<double> L(36 * (size_t)s.n_blocks, 0.0), A((size_t)N * N, 0.0), b(N), x(N);
  for (int j = 0; j < n; ++j)
    for (int bb = s.col_ptr[j]; bb < s.col_ptr[j + 1]; ++bb) {
      const int i = s.blk_row[bb];
      for (int r = 0; r < 6; ++r)
        for (int c = 0; c < 6; ++c) {
          double v = (i == j) ? (r == c ? 40.0 + U(rng) : 0.0) : 0.3 * U(rng);
          if (i == j && r != c) v = 0.0;
          L[36 * (size_t)bb + 6 * r + c] = v;
          A[(size_t)(6 * i + r) * N + 6 * j + c] = v;
          A[(size_t)(6 * j + c) * N + 6 * i + r] = v;
        }
    }
  for (int i = 0; i < N; ++i) b[i] = U(rng);
  const std::vector<double> L0 = L;
  // left-looking by levels: first every update task of the level, then factor its columns
  for (int lv = 0; lv < s.n_levels; ++lv) {
    CHECK(s.level_ptr[lv + 1] - s.level_ptr[lv] <= 64, "level too wide");
    if (s.ltask_ptr.empty()) {
      // no level tasks were planned (the tree program is in use): the same updates straight from the row lists
      for (int t = s.level_ptr[lv]; t < s.level_ptr[lv + 1]; ++t) {
        const int j = s.level_col[t];
        for (int rr = s.row_ptr[j]; rr < s.row_ptr[j + 1]; ++rr) {
          const int bjk = s.row_blk[rr], k = s.row_col[rr];
          for (int bik = bjk; bik < s.col_ptr[k + 1]; ++bik) {
            const int32_t *c0 = s.blk_row.data() + s.col_ptr[j], *c1 = s.blk_row.data() + s.col_ptr[j + 1];
            const int32_t *it = std::lower_bound(c0, c1, s.blk_row[bik]);
            CHECK(it != c1 && *it == s.blk_row[bik], "fill block missing");
            double *D = &L[36 * (size_t)(it - s.blk_row.data())];
            const double *Aa = &L[36 * (size_t)bik], *Bb = &L[36 * (size_t)bjk];
            for (int r = 0; r < 6; ++r) for (int c = 0; c < 6; ++c) { double sacc = 0; for (int kk = 0; kk < 6; ++kk) sacc += Aa[6 * r + kk] * Bb[6 * c + kk]; D[6 * r + c] -= sacc; }
          }
        }
      }
    }
    for (int t = s.ltask_ptr.empty() ? 0 : s.ltask_ptr[lv]; t < (s.ltask_ptr.empty() ? 0 : s.ltask_ptr[lv + 1]); ++t) {
      if (s.task_dst[t] < 0) { CHECK(s.level_col[s.level_ptr[lv] + s.task_pos[t]] == -1 - s.task_dst[t], "vec item pos"); continue; }
      CHECK(s.level_col[s.level_ptr[lv] + s.task_pos[t]] == s.blk_col[s.task_dst[t]], "item pos");
      double *D = &L[36 * (size_t)s.task_dst[t]];
      for (int u = s.task_pair_ptr[t]; u < s.task_pair_ptr[t + 1]; ++u) {
        const double *Aa = &L[36 * (size_t)s.pair_a[u]], *Bb = &L[36 * (size_t)s.pair_b[u]];
        CHECK(s.blk_col[s.pair_a[u]] == s.blk_col[s.pair_b[u]], "pair columns differ");
        CHECK(s.blk_row[s.pair_a[u]] == s.blk_row[s.task_dst[t]] && s.blk_row[s.pair_b[u]] == s.blk_col[s.task_dst[t]], "pair rows");
        for (int r = 0; r < 6; ++r) for (int c = 0; c < 6; ++c) { double sacc = 0; for (int k = 0; k < 6; ++k) sacc += Aa[6 * r + k] * Bb[6 * c + k]; D[6 * r + c] -= sacc; }
      }
    }
    for (int t = s.level_ptr[lv]; t < s.level_ptr[lv + 1]; ++t) {
      const int j = s.level_col[t];
      double *Djj = &L[36 * (size_t)s.col_ptr[j]];
      for (int jj = 0; jj < 6; ++jj) {
        double d = Djj[7 * jj];
        for (int k = 0; k < jj; ++k) d -= Djj[6 * jj + k] * Djj[6 * jj + k];
        CHECK(d > 0, "pivot");
        Djj[7 * jj] = std::sqrt(d);
        for (int i = jj + 1; i < 6; ++i) { double sacc = Djj[6 * i + jj]; for (int k = 0; k < jj; ++k) sacc -= Djj[6 * i + k] * Djj[6 * jj + k]; Djj[6 * i + jj] = sacc / Djj[7 * jj]; }
        for (int c = jj + 1; c < 6; ++c) Djj[6 * jj + c] = 0;
      }
      for (int bb = s.col_ptr[j] + 1; bb < s.col_ptr[j + 1]; ++bb)
        for (int r = 0; r < 6; ++r) {
          double *row = &L[36 * (size_t)bb + 6 * r], xr[6];
          for (int c = 0; c < 6; ++c) { double sacc = row[c]; for (int k = 0; k < c; ++k) sacc -= xr[k] * Djj[6 * c + k]; xr[c] = sacc / Djj[7 * c]; }
          for (int c = 0; c < 6; ++c) row[c] = xr[c];
        }
    }
  }
  for (int lv = 0; lv < s.n_levels; ++lv)
    for (int t = s.level_ptr[lv]; t < s.level_ptr[lv + 1]; ++t) {
      const int j = s.level_col[t];
      double sv[6];
      for (int r = 0; r < 6; ++r) {
        sv[r] = b[6 * j + r];
        for (int rr = s.row_ptr[j]; rr < s.row_ptr[j + 1]; ++rr) {
          CHECK(s.blk_row[s.row_blk[rr]] == j && s.blk_col[s.row_blk[rr]] == s.row_col[rr], "row list");
          for (int c = 0; c < 6; ++c) sv[r] -= L[36 * (size_t)s.row_blk[rr] + 6 * r + c] * x[6 * s.row_col[rr] + c];
        }
      }
      const double *Djj = &L[36 * (size_t)s.col_ptr[j]];
      for (int c = 0; c < 6; ++c) { double v = sv[c]; for (int k = 0; k < c; ++k) v -= Djj[6 * c + k] * x[6 * j + k]; x[6 * j + c] = v / Djj[7 * c]; }
    }
  for (int lv = s.n_levels - 1; lv >= 0; --lv)
    for (int t = s.level_ptr[lv]; t < s.level_ptr[lv + 1]; ++t) {
      const int j = s.level_col[t];
      double sv[6];
      for (int r = 0; r < 6; ++r) {
        sv[r] = x[6 * j + r];
        for (int bb = s.col_ptr[j] + 1; bb < s.col_ptr[j + 1]; ++bb)
          for (int c = 0; c < 6; ++c) sv[r] -= L[36 * (size_t)bb + 6 * c + r] * x[6 * s.blk_row[bb] + c];
      }
      const double *Djj = &L[36 * (size_t)s.col_ptr[j]];
      for (int c = 5; c >= 0; --c) { double v = sv[c]; for (int k = c + 1; k < 6; ++k) v -= Djj[6 * k + c] * x[6 * j + k]; x[6 * j + c] = v / Djj[7 * c]; }
    }
  // residual of the dense system
  double rmax = 0, bmax = 0;
  for (int i = 0; i < N; ++i) {
    double r = -b[i];
    for (int k = 0; k < N; ++k) r += A[(size_t)i * N + k] * x[k];
    rmax = std::fmax(rmax, std::fabs(r)); bmax = std::fmax(bmax, std::fabs(b[i]));
  }
  CHECK(rmax < 1e-10 * (1 + bmax), "residual %.3e (n=%d blocks=%d levels=%d)", rmax, n, s.n_blocks, s.n_levels);
  if (s.tree.ok) {
    const std::vector<double> xt = interpret_tree_program(s, L0, b);
    double dmax = 0;
    bool finite = true;
    for (int i = 0; i < N; ++i) { finite = finite && std::isfinite(xt[i]); dmax = std::fmax(dmax, std::fabs(xt[i] - x[i])); }
    CHECK(finite, "tree program: non-finite solution (a pool double was read before it was written?)");
    CHECK(dmax < 1e-10, "tree program: solution differs by %.3e (n=%d, cluster %d)", dmax, n, s.tree.C);
    std::printf("  tree program: C=%d top=%d chain=%d steps smem=%zu B words=%zu xchg=%d doubles\n", s.tree.C, s.tree.n_top_cols, s.tree.chain_steps,
                s.tree.smem_bytes, s.tree.words.size(), s.tree.xchg_doubles);
  } else {
    std::printf("  tree program not used: %s\n", s.tree.why_not.c_str());
    // the same system through the packed device program
    const std::vector<double> xp = interpret_program(s, L0, b);
    double dmax = 0;
    bool finite = true;
    for (int i = 0; i < N; ++i) { finite = finite && std::isfinite(xp[i]); dmax = std::fmax(dmax, std::fabs(xp[i] - x[i])); }
    CHECK(finite, "device program: non-finite solution (a block cache slot was read before it was written?)");
    CHECK(dmax < 1e-10, "device program: solution differs by %.3e (n=%d, cluster %d)", dmax, n, s.solve_cluster);
  }
  std::printf("case nk=%d np=%d w=%d fix0=%d nfix=%d loop=%d world=%d: n_fp=%d blocks=%d schur=%d levels=%d tasks=%d pairs=%d est=%.0f slots=%d cached=%d resid=%.2e\n",
              nk, np, w, (int)fix0, nfix, (int)loop, world, n, s.n_blocks, s.n_schur_blocks, s.n_levels, s.n_tasks, (int)s.pair_a.size(), s.est_solver_cycles, s.solver_slots, s.solver_cached_blocks, rmax);
}

// ---- pre-sharded input (ssba_options.presharded): every rank builds from the edges of its own landmarks; the ranks
// agree on the active sets and the co-visibility pattern through the callback.  Emulated with one thread per rank
// and a shared reducer; the result must give every rank the factor pattern of the single-rank build.
#include <condition_variable>
#include <mutex>
#include <thread>
struct Reducer {
  int world, arrived = 0, gen = 0;
  std::mutex mu; std::condition_variable cv;
  std::vector<uint8_t> bytes[2]; std::vector<long long> sums[2];  // by round parity: a round's result survives the next round's start
  bool failed = false;
  bool operator()(uint8_t *b, size_t nb, long long *s, int ns) {
    std::unique_lock<std::mutex> lk(mu);
    const int my = gen, k = my & 1;
    if (arrived == 0) { bytes[k].assign(nb, 0); sums[k].assign(ns, 0); }
    if (bytes[k].size() != nb || (int)sums[k].size() != ns) failed = true;  // (still take part: nobody may be left waiting)
    else {
      for (size_t i = 0; i < nb; ++i) bytes[k][i] = std::max(bytes[k][i], b[i]);
      for (int i = 0; i < ns; ++i) sums[k][i] += s[i];
    }
    if (++arrived == world) { arrived = 0; ++gen; cv.notify_all(); }
    else cv.wait(lk, [&] { return gen != my; });
    if (failed) return false;
    for (size_t i = 0; i < nb; ++i) b[i] = bytes[k][i];
    for (int i = 0; i < ns; ++i) s[i] = sums[k][i];
    return true;
  }
};
static void check_presharded(int nk, int np, int w, unsigned seed, bool fix0, int nfix, bool loop, int world) {
  HostGraph g = make_graph(nk, np, w, seed, fix0, nfix, loop);
  Structure full;
  std::string err;
  CHECK(build_structure(g, 0, 1, full, err), "full build: %s", err.c_str());
  std::vector<HostGraph> parts(world, g);
  for (int r = 0; r < world; ++r) {
    HostGraph &h = parts[r];
    h.e_pose.clear(); h.e_point.clear(); h.e_cam.clear(); h.e_uv.clear();
    for (int e = 0; e < g.n_edges; ++e)
      if (g.e_point[e] % world == r) { h.e_pose.push_back(g.e_pose[e]); h.e_point.push_back(g.e_point[e]); h.e_cam.push_back(g.e_cam[e]); h.e_uv.push_back(0); h.e_uv.push_back(0); }
    h.n_edges = (int)h.e_pose.size();
  }
  Reducer red; red.world = world;
  const AcrossRanks ar = [&](uint8_t *b, size_t nb, long long *s, int ns) { return red(b, nb, s, ns); };
  std::vector<Structure> S(world);
  std::vector<std::string> errs(world);
  std::vector<int> ok(world, 0);
  std::vector<std::thread> th;
  for (int r = 0; r < world; ++r) th.emplace_back([&, r] { ok[r] = build_structure(parts[r], r, world, S[r], errs[r], nullptr, &ar) ? 1 : 0; });
  for (auto &t : th) t.join();
  long long slots = 0, edges = 0;
  std::vector<int> owner(np, -1);
  for (int r = 0; r < world; ++r) {
    CHECK(ok[r], "pre-sharded build of rank %d: %s", r, errs[r].c_str());
    if (!ok[r]) return;
    const Structure &s = S[r];
    CHECK(s.n_fp == full.n_fp && s.n_blocks == full.n_blocks && s.col_ptr == full.col_ptr && s.blk_row == full.blk_row, "rank %d: factor pattern differs from the single-rank build", r);
    CHECK(s.pose_of_q == full.pose_of_q, "rank %d: elimination order differs", r);
    CHECK(s.n_active_edges_global == full.n_active_edges_global && s.n_fl_global == full.n_fl_global, "rank %d: global counts %d %d vs %d %d", r, s.n_active_edges_global, s.n_fl_global, full.n_active_edges_global, full.n_fl_global);
    CHECK(s.point_active == full.point_active, "rank %d: active landmarks differ", r);
    CHECK(s.tree.ok == full.tree.ok && s.tree.words == full.tree.words, "rank %d: solver program differs", r);
    for (int v : s.slot_vertex) { CHECK(owner[v] < 0, "landmark %d on two ranks", v); owner[v] = r; CHECK(v % world == r, "landmark %d on the wrong rank", v); }
    slots += s.n_slots; edges += s.n_edges;
    for (int c : s.combo_blk) CHECK(c >= 0 && c < s.n_blocks, "combo block out of range");
  }
  CHECK(slots == full.n_slots && edges == full.n_edges, "shards hold %lld slots / %lld edges of %d / %d", slots, edges, full.n_slots, full.n_edges);
}

static void check_block_inverse() {
  std::mt19937 rng(7);
  std::uniform_real_distribution<double> U(-1, 1);
  for (int trial = 0; trial < 2000; ++trial) {
    // D = G G^T + eps I with rows of very different scale (rotation / translation blocks of a pose)
    double G[36], D[36], Dm[36];
    for (int i = 0; i < 36; ++i) G[i] = U(rng) * ((i / 6) < 3 ? 1e3 : 1.0);
    for (int i = 0; i < 6; ++i) for (int k = 0; k < 6; ++k) { double v = 0; for (int m = 0; m < 6; ++m) v += G[6 * i + m] * G[6 * k + m]; D[6 * i + k] = v + (i == k ? 1e-3 : 0.0); }
    const int flip = trial % 4 == 3 ? (int)(rng() % 6) : -1;  // every fourth block is made indefinite
    if (flip >= 0) D[7 * flip] = -D[7 * flip];
    for (int i = 0; i < 36; ++i) Dm[i] = D[i];
    const bool bad = block_inverse6(Dm);
    // Cholesky verdict
    double L[36] = {0}; bool chol_bad = false;
    for (int j = 0; j < 6; ++j) {
      double dj = D[7 * j]; for (int k = 0; k < j; ++k) dj -= L[6 * j + k] * L[6 * j + k];
      if (!(dj > 0)) { chol_bad = true; break; }
      L[7 * j] = std::sqrt(dj);
      for (int i = j + 1; i < 6; ++i) { double v = D[6 * i + j]; for (int k = 0; k < j; ++k) v -= L[6 * i + k] * L[6 * j + k]; L[6 * i + j] = v / L[7 * j]; }
    }
    CHECK(bad == chol_bad, "block_inverse6: positive-definiteness verdict %d vs Cholesky %d (trial %d)", (int)bad, (int)chol_bad, trial);
    if (bad || chol_bad) continue;
    // residual D M - I, relative to the conditioning
    double res = 0, nm = 0, nd = 0;
    for (int i = 0; i < 36; ++i) { nm = std::max(nm, std::fabs(Dm[i])); nd = std::max(nd, std::fabs(D[i])); }
    for (int i = 0; i < 6; ++i) for (int k = 0; k < 6; ++k) { double v = 0; for (int m = 0; m < 6; ++m) v += D[6 * i + m] * Dm[6 * m + k]; res = std::max(res, std::fabs(v - (i == k ? 1.0 : 0.0))); }
    CHECK(res <= 1e-13 * nm * nd * 36, "block_inverse6: residual %g (|D| %g |M| %g)", res, nd, nm);
    for (int i = 0; i < 6; ++i) for (int k = 0; k < i; ++k) CHECK(Dm[6 * i + k] == Dm[6 * k + i], "block_inverse6: result not symmetric");
  }
}

int main() {
  check_block_inverse();
  check_presharded(30, 800, 5, 3, true, 25, false, 2);
  check_presharded(64, 3000, 5, 5, true, 0, true, 4);
  check_presharded(100, 4000, 5, 6, false, 0, false, 8);
  check_case(4, 40, 3, 1, false, 0, false, 1);
  check_case(10, 500, 3, 2, false, 0, false, 1);
  check_case(30, 800, 5, 3, true, 25, false, 2);
  check_case(40, 1500, 4, 4, false, 10, true, 3);
  check_case(64, 3000, 5, 5, true, 0, true, 8);
  check_case(100, 4000, 5, 6, false, 0, false, 1);
  check_case(500, 20000, 5, 7, true, 0, false, 1);
  check_case(12, 600, 12, 8, false, 0, false, 1);
  check_case(48, 1500, 6, 9, false, 0, true, 1);    // 32 <= poses < 64: the 4-CTA cluster
  check_case(33, 400, 33, 10, true, 0, false, 1);   // dense co-visibility: one separator, natural order
  check_case(150, 3000, 3, 11, false, 40, true, 4); // narrow band with long-range loops
  check_case(140, 6, 140, 12, false, 0, false, 2);  // landmarks seen by more poses than a CTA has threads: the big-chunk path
  if (fails) { std::printf("%d FAILURES\n", fails); return 1; }
  std::printf("OK\n");
  return 0;
}
