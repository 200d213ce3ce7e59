// Developer aid (CPU only): structure + k_tree_solve program statistics of a graph dumped by
// scripts/dev/tree_stats.py.  usage: tree_stats <graph.bin>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <vector>
#include "ssba_structure.hpp"
using namespace ssba;
int main(int argc, char **argv) {
  FILE *f = std::fopen(argv[1], "rb");
  int hdr[3];
  if (!f || std::fread(hdr, 4, 3, f) != 3) return 1;
  HostGraph g;
  g.have_cams = true; g.cams.n = 2;
  g.n_poses = hdr[0]; g.n_points = hdr[1]; g.n_edges = hdr[2];
  g.poses.assign(7 * (size_t)g.n_poses, 0.0); g.points.assign(3 * (size_t)g.n_points, 0.0);
  g.pose_fixed.resize(g.n_poses); g.point_fixed.resize(g.n_points);
  g.e_pose.resize(g.n_edges); g.e_point.resize(g.n_edges); g.e_cam.resize(g.n_edges); g.e_uv.resize(2 * (size_t)g.n_edges);
  size_t ok = std::fread(g.pose_fixed.data(), 1, g.n_poses, f) + std::fread(g.point_fixed.data(), 1, g.n_points, f) +
              std::fread(g.e_pose.data(), 4, g.n_edges, f) + std::fread(g.e_point.data(), 4, g.n_edges, f) + std::fread(g.e_cam.data(), 1, g.n_edges, f);
  (void)ok;
  std::fclose(f);
  Structure s;
  std::string err;
  if (!build_structure(g, 0, 1, s, err)) { std::printf("build failed: %s\n", err.c_str()); return 1; }
  std::printf("n_fp %d blocks %d schur blocks %d levels %d symbolic %.3f ms\n", s.n_fp, s.n_blocks, s.n_schur_blocks, s.n_levels, 1e3 * s.seconds_symbolic);
  {
    int mx = 0; long long tot = 0; std::vector<int> hist(12, 0);
    for (int b = 0; b < s.n_blocks; ++b) { const int n = s.blk_prod_ptr[b + 1] - s.blk_prod_ptr[b]; mx = std::max(mx, n); tot += n; int k = 0; while ((1 << k) < n + 1 && k < 11) ++k; ++hist[k]; }
    std::printf("schur units %d combos %zu  producers per block: max %d mean %.1f  histogram (<=1,2,4,...):", s.n_units, s.combo_blk.size(), mx, (double)tot / s.n_blocks);
    for (int k = 0; k < 12; ++k) std::printf(" %d", hist[k]);
    std::printf("\n");
  }
  {
    // fundamental / relaxed supernode chains in the final order: column j merges into its parent p when p == j + 1 is its
    // etree parent and struct(j) = struct(p) + {p} (fundamental), or with `extra` explicit zero blocks (relaxed)
    const int n = s.n_fp;
    int fund = 0, relax1 = 0, relax2 = 0;
    std::vector<int> chain(n, 1);
    for (int j = 0; j + 1 < n; ++j) {
      const int nbj = s.col_ptr[j + 1] - s.col_ptr[j] - 1;
      if (nbj == 0 || s.blk_row[s.col_ptr[j] + 1] != j + 1) continue;
      const int nbp = s.col_ptr[j + 2] - s.col_ptr[j + 1] - 1;
      const int extra = nbp + 1 - nbj;
      if (extra == 0) ++fund; else if (extra == 1) ++relax1; else if (extra == 2) ++relax2;
      std::printf("col %3d -> %3d: blocks %2d parent blocks %2d extra %d\n", j, j + 1, nbj, nbp, extra);
    }
    std::printf("parent == j+1 pairs: fundamental %d, one extra block %d, two %d (of %d columns)\n", fund, relax1, relax2, n);
  }
  const TreeProgram &tp = s.tree;
  if (!tp.ok) { std::printf("tree program not built: %s\n", tp.why_not.c_str()); return 0; }
  std::printf("tree: C %d smem %zu chain_steps %d top cols %d xchg doubles %d words %zu\n", tp.C, tp.smem_bytes, tp.chain_steps, tp.n_top_cols, tp.xchg_doubles, tp.words.size());
  for (int c = 0; c < tp.C; ++c) {
    const int32_t *w = tp.words.data() + tp.prog_ptr[c];
    const int nsa = w[kTH_StepsA], nsb = w[kTH_StepsB];
    std::printf("CTA %2d: cols %3d blocks %4d pool %5d doubles contrib %5d  stepsA %d stepsB %d  add rounds %d\n", c, tp.n_own_cols[c], tp.n_own_blocks[c], tp.pool_doubles[c], tp.contrib_doubles[c], nsa, nsb, w[kTH_AddRounds]);
    const int32_t *steps = w + w[kTH_OffSteps];
    long long tot_look = 0, tot_crit = 0;
    for (int st = 0; st < nsa + nsb; ++st) {
      const int32_t *e = steps + kTS_Words * st;
      int maxcrit = 0, sumcrit = 0;
      for (int t = 0; t < (e[kTS_Cols] & 0xffff); ++t) { const int np = (int)((unsigned)w[e[kTS_OffDiag] + 2 * t] >> 20); maxcrit = std::max(maxcrit, np); sumcrit += np; }
      long long lookp = 0; int maxlook = 0;
      for (int i = 0; i < 5 * e[kTS_NLook]; ++i) { const unsigned x = (unsigned)w[e[kTS_OffLook] + 2 * i]; const int np = x >> 20, nr = (x >> 16) & 15; lookp += (long long)np * nr; maxlook = std::max(maxlook, np); }
      tot_look += lookp; tot_crit += sumcrit;
      if (c == 0 || c == 1) std::printf("   step %2d%s: cols %2d crit pairs max %2d sum %3d | look rounds %3d (row-products %5lld, longest %2d) | panel rounds %3d\n", st, st >= nsa ? "T" : " ", (e[kTS_Cols] & 0xffff), maxcrit, sumcrit, e[kTS_NLook], lookp, maxlook, e[kTS_NPanel]);
    }
    std::printf("   total: critical block products %lld, look-ahead row products %lld (= %lld block products)\n", tot_crit, tot_look, tot_look / 6);
  }
  return 0;
}
