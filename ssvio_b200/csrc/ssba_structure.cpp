// ssba_structure.cpp — see ssba_structure.hpp.
#include "ssba_structure.hpp"

#include <algorithm>
#include <climits>
#include <cstring>
#include <numeric>

namespace ssba {

namespace {

constexpr int kHppChunk = 256;  // edges per pose-major chunk (one CTA each)

// Order of elimination of the free poses in the reduced system.  The reference runs AMD on the
// block pattern (linear_solver_csparse.h:262); any symmetric permutation gives the same x up to
// rounding.  Sliding-window graphs are block-banded in key-frame order, where the natural order
// is already fill-minimal, so that is what is used for now.
void order_free_poses(int n, const std::vector<std::vector<int>> & /*adj*/, std::vector<int> &perm) {
  perm.resize(n);
  std::iota(perm.begin(), perm.end(), 0);
}

}  // namespace

bool build_structure(const HostGraph &g, int rank, int world, Structure &s, std::string &err) {
  s = Structure();
  const int NK = g.n_poses, NP = g.n_points, NE = g.n_edges;
  s.n_edges_total = NE;
  if (!g.have_cams) { err = "ssba_set_cameras was not called"; return false; }
  for (int e = 0; e < NE; ++e) {
    if (g.e_pose[e] < 0 || g.e_pose[e] >= NK) { err = "edge pose index out of range"; return false; }
    if (g.e_point[e] < 0 || g.e_point[e] >= NP) { err = "edge point index out of range"; return false; }
    if (g.e_cam[e] >= g.cams.n) { err = "edge camera index out of range"; return false; }
  }

  // ---- active sets (sparse_optimizer.cpp:201-272): an edge is active unless both ends are fixed
  std::vector<uint8_t> pose_active(NK, 0), point_active(NP, 0), edge_active(NE, 0);
  std::vector<int32_t> point_deg(NP + 1, 0);
  int n_active = 0;
  for (int e = 0; e < NE; ++e) {
    if (g.pose_fixed[g.e_pose[e]] && g.point_fixed[g.e_point[e]]) continue;
    edge_active[e] = 1; ++n_active;
    pose_active[g.e_pose[e]] = 1; point_active[g.e_point[e]] = 1;
    ++point_deg[g.e_point[e]];
  }
  s.n_active_edges_global = n_active;

  // ---- index mapping (sparse_optimizer.cpp:168-192): free poses in id order, then landmarks
  std::vector<int32_t> fp_of_pose(NK, -1), free_pose_rows;
  for (int i = 0; i < NK; ++i)
    if (pose_active[i] && !g.pose_fixed[i]) { fp_of_pose[i] = (int)free_pose_rows.size(); free_pose_rows.push_back(i); }
  s.n_fp = (int)free_pose_rows.size();
  s.n_fl_global = 0;
  for (int j = 0; j < NP; ++j)
    if (point_active[j] && !g.point_fixed[j]) ++s.n_fl_global;

  // ---- edges by landmark (CSR over point rows, stable = addEdge order inside a landmark)
  std::vector<int32_t> pt_ptr(NP + 1, 0);
  for (int j = 0; j < NP; ++j) pt_ptr[j + 1] = pt_ptr[j] + point_deg[j];
  std::vector<int32_t> pt_edges(n_active);
  {
    std::vector<int32_t> fill(pt_ptr.begin(), pt_ptr.end() - 1);
    for (int e = 0; e < NE; ++e)
      if (edge_active[e]) pt_edges[fill[g.e_point[e]]++] = e;
  }

  // ---- co-visibility of free poses through free landmarks = pattern of the Schur complement
  // (block_solver.hpp:224-249), as a bitmap when small enough, else as a key list
  const int n = s.n_fp;
  const bool use_bitmap = n <= 8192;
  std::vector<uint64_t> bitmap;
  std::vector<uint64_t> keys;
  if (use_bitmap) bitmap.assign(((size_t)n * n + 63) / 64, 0);
  std::vector<int32_t> tmp;
  auto mark = [&](int r, int c) {  // r >= c, unpermuted free-pose indices
    const size_t bit = (size_t)c * n + r;
    if (use_bitmap) bitmap[bit >> 6] |= 1ull << (bit & 63); else keys.push_back(bit);
  };
  for (int i = 0; i < n; ++i) mark(i, i);
  for (int j = 0; j < NP; ++j) {
    if (!point_active[j] || g.point_fixed[j]) continue;
    tmp.clear();
    for (int k = pt_ptr[j]; k < pt_ptr[j + 1]; ++k) {
      const int f = fp_of_pose[g.e_pose[pt_edges[k]]];
      if (f >= 0) tmp.push_back(f);
    }
    std::sort(tmp.begin(), tmp.end());
    tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
    for (size_t a = 0; a < tmp.size(); ++a)
      for (size_t b = a + 1; b < tmp.size(); ++b) mark(tmp[b], tmp[a]);
  }
  std::vector<std::vector<int>> adj(n);  // strictly-lower rows per column, unpermuted
  if (use_bitmap) {
    for (int c = 0; c < n; ++c)
      for (int r = c + 1; r < n; ++r) {
        const size_t bit = (size_t)c * n + r;
        if (bitmap[bit >> 6] >> (bit & 63) & 1) adj[c].push_back(r);
      }
  } else {
    std::sort(keys.begin(), keys.end());
    keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
    for (uint64_t k : keys) { const int c = (int)(k / n), r = (int)(k % n); if (r != c) adj[c].push_back(r); }
  }

  // ---- elimination order and symbolic factorisation over q
  std::vector<int> perm;   // q -> free pose index
  order_free_poses(n, adj, perm);
  std::vector<int> iperm(n);
  for (int q = 0; q < n; ++q) iperm[perm[q]] = q;
  s.q_of_pose.assign(NK, -1);
  s.pose_of_q.resize(n);
  for (int f = 0; f < n; ++f) { s.q_of_pose[free_pose_rows[f]] = iperm[f]; s.pose_of_q[iperm[f]] = free_pose_rows[f]; }

  std::vector<std::vector<int>> Acol(n);  // permuted strictly-lower pattern of S
  int n_schur = n;
  for (int c = 0; c < n; ++c)
    for (int r : adj[c]) {
      int qc = iperm[c], qr = iperm[r];
      if (qr < qc) std::swap(qr, qc);
      Acol[qc].push_back(qr);
      ++n_schur;
    }
  s.n_schur_blocks = n_schur;
  // column structure of L: struct(L_j) = struct(A_j) U (U over children c: struct(L_c) \ {j})
  std::vector<std::vector<int>> Lcol(n);
  {
    std::vector<std::vector<int>> children(n);
    std::vector<int> mark_v(n, -1), merged;
    for (int j = 0; j < n; ++j) {
      merged.clear();
      mark_v[j] = j;
      for (int r : Acol[j]) if (mark_v[r] != j) { mark_v[r] = j; merged.push_back(r); }
      for (int c : children[j])
        for (int r : Lcol[c]) if (r != j && mark_v[r] != j) { mark_v[r] = j; merged.push_back(r); }
      std::sort(merged.begin(), merged.end());
      Lcol[j] = merged;
      if (!merged.empty()) children[merged[0]].push_back(j);  // etree parent = first sub-diagonal row
    }
  }
  s.col_ptr.assign(n + 1, 0);
  for (int j = 0; j < n; ++j) s.col_ptr[j + 1] = s.col_ptr[j] + 1 + (int)Lcol[j].size();
  s.n_blocks = s.col_ptr[n];
  s.blk_row.resize(s.n_blocks);
  s.blk_col.resize(s.n_blocks);
  for (int j = 0; j < n; ++j) {
    int b = s.col_ptr[j];
    s.blk_col[b] = j; s.blk_row[b++] = j;
    for (int r : Lcol[j]) { s.blk_col[b] = j; s.blk_row[b++] = r; }
  }
  auto find_block = [&](int row, int col) -> int {  // row >= col
    const int *b0 = s.blk_row.data() + s.col_ptr[col], *b1 = s.blk_row.data() + s.col_ptr[col + 1];
    const int *it = std::lower_bound(b0, b1, row);
    return (it != b1 && *it == row) ? (int)(it - s.blk_row.data()) : -1;
  };
  // strictly-lower blocks by row, ordered by column
  s.row_ptr.assign(n + 1, 0);
  for (int j = 0; j < n; ++j)
    for (int r : Lcol[j]) ++s.row_ptr[r + 1];
  for (int j = 0; j < n; ++j) s.row_ptr[j + 1] += s.row_ptr[j];
  s.row_blk.resize(s.row_ptr[n]);
  s.row_col.resize(s.row_ptr[n]);
  {
    std::vector<int32_t> fill(s.row_ptr.begin(), s.row_ptr.end() - 1);
    for (int j = 0; j < n; ++j)
      for (int b = s.col_ptr[j] + 1; b < s.col_ptr[j + 1]; ++b) {
        const int r = s.blk_row[b];
        s.row_blk[fill[r]] = b; s.row_col[fill[r]] = j; ++fill[r];
      }
  }
  // left-looking update lists: column j gathers L(i,k) L(j,k)^T for every k < j with L(j,k) != 0
  s.upd_ptr.assign(n + 1, 0);
  for (int j = 0; j < n; ++j) {
    int cnt = 0;
    for (int t = s.row_ptr[j]; t < s.row_ptr[j + 1]; ++t) {
      const int bjk = s.row_blk[t], k = s.row_col[t];
      cnt += s.col_ptr[k + 1] - bjk;  // rows i >= j of column k
    }
    s.upd_ptr[j + 1] = s.upd_ptr[j] + cnt;
  }
  s.upd_dst.resize(s.upd_ptr[n]); s.upd_a.resize(s.upd_ptr[n]); s.upd_b.resize(s.upd_ptr[n]);
  for (int j = 0; j < n; ++j) {
    int w = s.upd_ptr[j];
    for (int t = s.row_ptr[j]; t < s.row_ptr[j + 1]; ++t) {
      const int bjk = s.row_blk[t], k = s.row_col[t];
      for (int bik = bjk; bik < s.col_ptr[k + 1]; ++bik) {
        const int dst = find_block(s.blk_row[bik], j);
        if (dst < 0) { err = "internal: symbolic factorisation inconsistent"; return false; }
        s.upd_dst[w] = dst; s.upd_a[w] = bik; s.upd_b[w] = bjk; ++w;
      }
    }
  }
  // elimination-tree levels: column j can start once every k with L(j,k) != 0 is done
  {
    std::vector<int> level(n, 0);
    int nl = 0;
    for (int j = 0; j < n; ++j) {
      int lv = 0;
      for (int t = s.row_ptr[j]; t < s.row_ptr[j + 1]; ++t) lv = std::max(lv, level[s.row_col[t]] + 1);
      level[j] = lv; nl = std::max(nl, lv + 1);
    }
    s.n_levels = n ? nl : 0;
    s.level_ptr.assign(s.n_levels + 1, 0);
    for (int j = 0; j < n; ++j) ++s.level_ptr[level[j] + 1];
    for (int l = 0; l < s.n_levels; ++l) s.level_ptr[l + 1] += s.level_ptr[l];
    s.level_col.resize(n);
    std::vector<int32_t> fill(s.level_ptr.begin(), s.level_ptr.end() - 1);
    for (int j = 0; j < n; ++j) s.level_col[fill[level[j]]++] = j;
  }

  // ---- landmark shard of this rank: contiguous runs of active landmarks, balanced by edges
  std::vector<int32_t> slots;
  {
    const long long total = n_active;
    long long seen = 0;
    for (int j = 0; j < NP; ++j) {
      if (!point_active[j]) continue;
      // owner = the rank whose edge-quantile holds the first edge of this landmark
      const int owner = total > 0 ? (int)std::min<long long>(world - 1, seen * world / total) : 0;
      if (owner == rank) slots.push_back(j);
      seen += point_deg[j];
    }
  }
  s.n_slots = (int)slots.size();
  s.slot_vertex.assign(slots.begin(), slots.end());
  s.slot_free.resize(s.n_slots);
  s.slot_pair_ptr.assign(s.n_slots + 1, 0);
  s.slot_combo_ptr.assign(s.n_slots + 1, 0);
  const bool have_info = !g.e_info.empty(), have_delta = !g.e_delta.empty();
  std::vector<std::pair<long long, int32_t>> order;  // (pair key, edge)
  s.pair_edge_ptr.push_back(0);
  std::vector<int32_t> wq;
  for (int sl = 0; sl < s.n_slots; ++sl) {
    const int j = slots[sl];
    const bool lfree = !g.point_fixed[j];
    s.slot_free[sl] = lfree;
    if (lfree) ++s.n_fl;
    order.clear();
    for (int k = pt_ptr[j]; k < pt_ptr[j + 1]; ++k) {
      const int e = pt_edges[k];
      const int q = s.q_of_pose[g.e_pose[e]];
      // pairs with a free pose first, by q; fixed-pose pairs after, by pose row
      const long long key = q >= 0 ? q : (long long)n + g.e_pose[e];
      order.emplace_back(key, e);
    }
    std::sort(order.begin(), order.end());
    wq.clear();
    long long last = -1;
    for (auto &ke : order) {
      const int e = ke.second;
      if (ke.first != last) {
        if (last != -1) s.pair_edge_ptr.push_back((int32_t)s.e_orig.size());
        last = ke.first;
        s.pair_vertex.push_back(g.e_pose[e]);
        const int q = s.q_of_pose[g.e_pose[e]];
        s.pair_q.push_back(q);
        if (q >= 0 && lfree) wq.push_back(q);
      }
      s.e_orig.push_back(e);
      s.e_uv.push_back(g.e_uv[2 * e]); s.e_uv.push_back(g.e_uv[2 * e + 1]);
      s.e_cam.push_back(g.e_cam[e]);
      if (have_info) for (int t = 0; t < 3; ++t) s.e_info.push_back(g.e_info[3 * e + t]);
      if (have_delta) s.e_delta.push_back(g.e_delta[e]);
    }
    if (last != -1) s.pair_edge_ptr.push_back((int32_t)s.e_orig.size());
    s.slot_pair_ptr[sl + 1] = (int32_t)s.pair_vertex.size();
    // Schur targets: for W-pairs a <= b (sorted by q): block (row q_b, col q_a)
    for (size_t a = 0; a < wq.size(); ++a)
      for (size_t b = a; b < wq.size(); ++b) {
        const int blk = find_block(wq[b], wq[a]);
        if (blk < 0) { err = "internal: Schur block missing from the factor pattern"; return false; }
        s.combo_blk.push_back(blk);
      }
    s.slot_combo_ptr[sl + 1] = (int32_t)s.combo_blk.size();
  }
  s.n_pairs = (int)s.pair_vertex.size();
  s.n_edges = (int)s.e_orig.size();

  // ---- pose-major copy (edges of this shard whose pose is free), cut into one-pose chunks
  {
    std::vector<int32_t> cnt(n + 1, 0);
    for (int a = 0; a < s.n_pairs; ++a)
      if (s.pair_q[a] >= 0) cnt[s.pair_q[a] + 1] += s.pair_edge_ptr[a + 1] - s.pair_edge_ptr[a];
    for (int q = 0; q < n; ++q) cnt[q + 1] += cnt[q];
    s.n_pm_edges = cnt[n];
    s.pm_uv.resize(2 * (size_t)s.n_pm_edges); s.pm_point.resize(s.n_pm_edges); s.pm_cam.resize(s.n_pm_edges);
    if (have_info) s.pm_info.resize(3 * (size_t)s.n_pm_edges);
    if (have_delta) s.pm_delta.resize(s.n_pm_edges);
    std::vector<int32_t> fill(cnt.begin(), cnt.end() - 1);
    for (int sl = 0; sl < s.n_slots; ++sl)
      for (int a = s.slot_pair_ptr[sl]; a < s.slot_pair_ptr[sl + 1]; ++a) {
        const int q = s.pair_q[a];
        if (q < 0) continue;
        for (int e = s.pair_edge_ptr[a]; e < s.pair_edge_ptr[a + 1]; ++e) {
          const int d = fill[q]++;
          s.pm_uv[2 * d] = s.e_uv[2 * e]; s.pm_uv[2 * d + 1] = s.e_uv[2 * e + 1];
          s.pm_point[d] = s.slot_vertex[sl]; s.pm_cam[d] = s.e_cam[e];
          if (have_info) for (int t = 0; t < 3; ++t) s.pm_info[3 * d + t] = s.e_info[3 * e + t];
          if (have_delta) s.pm_delta[d] = s.e_delta[e];
        }
      }
    s.q_chunk_ptr.assign(n + 1, 0);
    s.chunk_edge_ptr.push_back(0);
    for (int q = 0; q < n; ++q) {
      for (int e0 = cnt[q]; e0 < cnt[q + 1]; e0 += kHppChunk) {
        s.chunk_q.push_back(q);
        s.chunk_vertex.push_back(s.pose_of_q[q]);
        s.chunk_edge_ptr.push_back(std::min(e0 + kHppChunk, cnt[q + 1]));
      }
      s.q_chunk_ptr[q + 1] = (int32_t)s.chunk_q.size();
    }
    s.n_chunks = (int)s.chunk_q.size();
  }
  return true;
}

}  // namespace ssba
