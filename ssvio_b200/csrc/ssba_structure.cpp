// ssba_structure.cpp — see ssba_structure.hpp.
#include "ssba_structure.hpp"
#include "ssba_solver_layout.hpp"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <condition_variable>
#include <cstdlib>
#include <functional>
#include <mutex>
#include <numeric>
#include <thread>

namespace ssba {

static std::atomic<int> g_cluster_cap{8};
void set_solver_cluster_cap(int cap) { g_cluster_cap.store(cap >= 8 ? 8 : cap >= 4 ? 4 : cap >= 2 ? 2 : 1); }

int solver_cluster_size(int n_fp) {
  static const int configured = [] {
    const char *e = std::getenv("SSBA_SOLVE_CLUSTER");
    const int c = e ? std::atoi(e) : 0;
    return (c == 1 || c == 2 || c == 4 || c == 8) ? c : 0;
  }();
  if (n_fp < 32) return 1;
  const int want = configured ? configured : (n_fp < 64 ? 4 : 8);
  return std::min(want, g_cluster_cap.load());  // measured on B200: cfg3 (100 poses) 4 -> 8 CTAs: -2 %, cfg5 (500 poses): -13 %
}

namespace {

constexpr int kMaxLevelCols = kSolveMaxCols;  // ssba_solver_layout.hpp
constexpr int kSchurRunPairs = 80;  // = kSchurRunPairs of k_schur
constexpr int kSchurRun = 16;       // = kSchurRun of k_schur
constexpr int kLinPairs = 128;      // = kLinThreads of k_linearize / k_update

// Symbolic factorisation of the reduced system under one elimination order, plus the schedule
// the device solver walks: columns grouped by elimination-tree level, and per level the
// left-looking update tasks  L(i,j) -= sum_k L(i,k) L(j,k)^T  grouped by destination block.
struct Factor {
  int n_blocks = 0, n_levels = 0, n_schur = 0;
  double est_cycles = 0.0;
  std::vector<int32_t> col_ptr, blk_row, blk_col, row_ptr, row_blk, row_col, level_ptr, level_col;
  std::vector<int32_t> ltask_ptr, task_dst, task_pos, task_pair_ptr, pair_a, pair_b;
};

bool symbolic_factor(int n, const std::vector<std::vector<int>> &adj, const std::vector<int> &perm,
                     Factor &f, std::string &err, bool want_tasks = true) {
  std::vector<int> iperm(n);
  for (int q = 0; q < n; ++q) iperm[perm[q]] = q;
  std::vector<std::vector<int>> Acol(n);  // permuted strictly-lower pattern of S
  f.n_schur = n;
  for (int c = 0; c < n; ++c)
    for (int r : adj[c]) {
      int qc = iperm[c], qr = iperm[r];
      if (qr < qc) std::swap(qr, qc);
      Acol[qc].push_back(qr);
      ++f.n_schur;
    }
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
  f.col_ptr.assign(n + 1, 0);
  for (int j = 0; j < n; ++j) f.col_ptr[j + 1] = f.col_ptr[j] + 1 + (int)Lcol[j].size();
  f.n_blocks = f.col_ptr[n];
  f.blk_row.resize(f.n_blocks);
  f.blk_col.resize(f.n_blocks);
  for (int j = 0; j < n; ++j) {
    int b = f.col_ptr[j];
    f.blk_col[b] = j; f.blk_row[b++] = j;
    for (int r : Lcol[j]) { f.blk_col[b] = j; f.blk_row[b++] = r; }
  }
  auto find_block = [&](int row, int col) -> int {
    const int *b0 = f.blk_row.data() + f.col_ptr[col], *b1 = f.blk_row.data() + f.col_ptr[col + 1];
    const int *it = std::lower_bound(b0, b1, row);
    return (it != b1 && *it == row) ? (int)(it - f.blk_row.data()) : -1;
  };
  // strictly-lower blocks by row, ordered by column
  f.row_ptr.assign(n + 1, 0);
  for (int j = 0; j < n; ++j)
    for (int r : Lcol[j]) ++f.row_ptr[r + 1];
  for (int j = 0; j < n; ++j) f.row_ptr[j + 1] += f.row_ptr[j];
  f.row_blk.resize(f.row_ptr[n]);
  f.row_col.resize(f.row_ptr[n]);
  {
    std::vector<int32_t> fill(f.row_ptr.begin(), f.row_ptr.end() - 1);
    for (int j = 0; j < n; ++j)
      for (int b = f.col_ptr[j] + 1; b < f.col_ptr[j + 1]; ++b) {
        const int r = f.blk_row[b];
        f.row_blk[fill[r]] = b; f.row_col[fill[r]] = j; ++fill[r];
      }
  }
  // elimination-tree levels: column j can start once every k with L(j,k) != 0 is done; a level
  // holds at most kMaxLevelCols columns (the device keeps one inverse diagonal block per
  // column of the running level in shared memory), wider levels are cut into several
  std::vector<int> level(n, 0);
  {
    std::vector<int> dep(n, 0);
    int nl = 0;
    for (int j = 0; j < n; ++j) {
      int lv = 0;
      for (int t = f.row_ptr[j]; t < f.row_ptr[j + 1]; ++t) lv = std::max(lv, dep[f.row_col[t]] + 1);
      dep[j] = lv; nl = std::max(nl, lv + 1);
    }
    std::vector<std::vector<int>> by_dep(nl);
    for (int j = 0; j < n; ++j) by_dep[dep[j]].push_back(j);
    f.level_ptr.assign(1, 0);
    f.level_col.clear();
    for (int d = 0; d < nl; ++d)
      for (size_t o = 0; o < by_dep[d].size(); o += kMaxLevelCols) {
        const size_t e = std::min(by_dep[d].size(), o + kMaxLevelCols);
        for (size_t i = o; i < e; ++i) { level[by_dep[d][i]] = (int)f.level_ptr.size() - 1; f.level_col.push_back(by_dep[d][i]); }
        f.level_ptr.push_back((int32_t)f.level_col.size());
      }
    f.n_levels = (int)f.level_ptr.size() - 1;
  }
  if (!want_tasks) return true;  // the subtree-per-CTA program (ssba_tree_program.cpp) plans its own work items
  // Work items of the device solver, level by level.  Every block of a column of the level is
  // one item:  L(i,j) <- (A(i,j) - sum_k L(i,k) L(j,k)^T) * L(j,j)^-T, the diagonal items of
  // all columns first (the others wait for the inverse diagonal block they produce), then one
  // "vector" item per column for the fused forward substitution y_j.  The pairs of an item are
  // ordered by source column k (deterministic summation order).
  f.ltask_ptr.assign(f.n_levels + 1, 0);
  f.task_pair_ptr.assign(1, 0);
  std::vector<std::vector<std::vector<std::pair<int, int>>>> per_dst;  // [col in level][block][pairs]
  f.est_cycles = 0.0;
  for (int lv = 0; lv < f.n_levels; ++lv) {
    const int c0 = f.level_ptr[lv], nc = f.level_ptr[lv + 1] - c0;
    per_dst.assign(nc, {});
    long long level_pairs = 0; int max_rounds = 0, n_items = 0;
    for (int t = 0; t < nc; ++t) {
      const int j = f.level_col[c0 + t];
      per_dst[t].assign(f.col_ptr[j + 1] - f.col_ptr[j], {});
      for (int rr = f.row_ptr[j]; rr < f.row_ptr[j + 1]; ++rr) {
        const int bjk = f.row_blk[rr], k = f.row_col[rr];
        for (int bik = bjk; bik < f.col_ptr[k + 1]; ++bik) {  // rows i >= j of column k
          const int dst = find_block(f.blk_row[bik], j);
          if (dst < 0) { err = "internal: symbolic factorisation inconsistent"; return false; }
          per_dst[t][dst - f.col_ptr[j]].emplace_back(bik, bjk);
        }
      }
    }
    auto emit = [&](int t, int d) {
      const int j = f.level_col[c0 + t];
      f.task_dst.push_back(f.col_ptr[j] + d);
      f.task_pos.push_back(t);
      for (auto &ab : per_dst[t][d]) { f.pair_a.push_back(ab.first); f.pair_b.push_back(ab.second); }
      f.task_pair_ptr.push_back((int32_t)f.pair_a.size());
      level_pairs += (long long)per_dst[t][d].size();
      max_rounds = std::max(max_rounds, ((int)per_dst[t][d].size() + 4) / 5);
      ++n_items;
    };
    for (int t = 0; t < nc; ++t) emit(t, 0);                       // diagonal items first
    for (int t = 0; t < nc; ++t)
      for (int d = 1; d < (int)per_dst[t].size(); ++d) emit(t, d);   // sub-diagonal items
    for (int t = 0; t < nc; ++t) {                                  // forward-substitution items
      f.task_dst.push_back(-1 - f.level_col[c0 + t]);
      f.task_pos.push_back(t);
      f.task_pair_ptr.push_back((int32_t)f.pair_a.size());
      ++n_items;
    }
    f.ltask_ptr[lv + 1] = (int32_t)f.task_dst.size();
    // cost model of one level on one SM (cycles): items round-robin over 32 warps, each paying
    // an L2 round trip, plus the diagonal factor on the critical path and a barrier
    const double rounds_per_warp = std::ceil(n_items / 32.0);
    f.est_cycles += std::max(900.0 * rounds_per_warp + 100.0 * max_rounds, 5.0 * (double)level_pairs) + 1400.0;
  }
  f.est_cycles += 1200.0 * f.n_levels;  // backward substitution
  return true;
}

// Packs everything the device solver needs to walk one level into one contiguous int32 segment
// (a level's indices are staged into shared memory with one asynchronous copy).  Segment 0 is a
// prologue without columns that only prefetches the blocks of level 0; segment 1 + lv is level lv.
//
// The work of a level is cut into warp "rounds" of five group tasks (a warp = 5 groups of 6
// lanes, lane r of a group owns row r of a 6x6 block):
//   DIAG round  five diagonal blocks: update, 6x6 Cholesky, inverse -> published per column
//   SUB  round  five sub-diagonal blocks: update, wait for the column's inverse, multiply
//   VEC  round  five forward substitutions y_j
// A block with many update pairs gets a round of its own with its pairs split five ways
// (REDUCE flag: the groups' partial sums are added before group 0 finishes the block).
//   header[8] = {n_cols, n_rounds, n_pairs, n_brows, n_pf, n_bpf, 0, 0}
//   cta_rptr[kSolveMaxCluster + 1]             rounds [cta_rptr[c], cta_rptr[c+1]) belong to CTA c of the cluster
//   col_j[n_cols] col_b0[n_cols] col_bptr[n_cols+1] brow[n_brows]          (backward pass)
//   round_type[n_rounds]                       bits 0-1: 0 DIAG 1 SUB 2 VEC 3 ACC (look-ahead), bit 2: REDUCE
//   gt_dst[5 n_rounds] gt_slot[..] gt_pos[..] gt_p0[..] gt_p1[..] gt_mask[..]   (group tasks; mask = CTAs that read the SUB block later)
//   pa[n_pairs] pb[n_pairs]      block refs; for VEC tasks pb is the column k of y_k
//   pf_blk[n_pf] pf_slot[n_pf]   blocks of the NEXT level to prefetch into their cache slots
//   bpf_blk[n_bpf]               blocks of the PREVIOUS level (backward pass prefetch)
// Factor blocks live in global memory; a shared-memory cache of `n_slots` 6x6 slots is planned
// here on the host (interval allocation over the levels).  A block reference r >= 0 is slot r,
// r < 0 is block -1-r in global memory.
void build_solver_program(Structure &s) {
  constexpr int kSplitPairs = 10;  // blocks with more update pairs get a REDUCE round
  constexpr int kSplitDiag = 1;    // ... diagonal blocks already with two
  // (measured on B200, cfg3: the solve time moves by < 1 % for kSplitAcc 3..10, kAccMinPairs 1..6 and a
  // level capacity of 100..250 % of the even share)
  constexpr int kSplitAcc = 5;     // ... look-ahead tasks with more than five (they must not outlast the level's chain)
  constexpr int kAccMinPairs = 3;  // tasks with more update pairs hand the early ones to the levels before
  constexpr int kAccCapPct = 150;  // look-ahead work a level takes, in % of an even share of all products  // tasks with more update pairs hand the early ones to the level before
  const int n = s.n_fp, NL = s.n_levels;
  std::vector<int> level_of(n, 0);
  for (int lv = 0; lv < NL; ++lv)
    for (int t = s.level_ptr[lv]; t < s.level_ptr[lv + 1]; ++t) level_of[s.level_col[t]] = lv;
  auto level_blocks = [&](int lv) {
    int nb = 0;
    for (int t = s.level_ptr[lv]; t < s.level_ptr[lv + 1]; ++t) { const int j = s.level_col[t]; nb += s.col_ptr[j + 1] - s.col_ptr[j]; }
    return nb;
  };
  // ---- rounds of every level (slot-independent part)
  struct GT { int dst, pos, p0, p1; };             // pairs index the level-local pair arrays
  struct Round { int type; GT g[5]; };
  struct LevelPlan { std::vector<Round> rounds; std::vector<std::pair<int, int>> pairs; /* (a block, b block | column) */ int cta_rptr[kSolveMaxCluster + 1]; };
  std::vector<LevelPlan> plan(NL);
  const int C = solver_cluster_size(n);
  s.solve_cluster = C;
  // ---- look-ahead: a task of level lv with more than kAccMinPairs update products hands the ones
  // whose source column finished at a level sl <= lv - 2 to ACC tasks of earlier levels in
  // [sl + 1, lv - 1] (which pre-subtract them from the block's value in global memory on otherwise
  // idle warps / CTAs); at its own level only the products of the columns of level lv - 1 are left
  // on the critical path.  Measured on B200 (cfg3): everything at lv - 1 overloads the levels below
  // the top separators, everything at sl + 1 the leaf levels; hence the load-levelled placement.
  struct Task { int dst, pos, begin, end; };  // its pairs are tpairs[begin, end); VEC: dst = -1 - column, pairs (block, -1 - col)
  std::vector<std::pair<int, int>> tpairs;
  tpairs.reserve(s.pair_a.size() + s.row_blk.size() + 16);
  std::vector<std::vector<Task>> own(NL), early(NL);
  {
    // pass 1: the late products stay with their task; count the work every level has anyway
    std::vector<long long> load(NL, 0);
    long long total_pairs = 0;
    struct Group { int lv, dst, sl, begin, end; };  // early products of one block from the columns of one level: gpairs[begin, end)
    std::vector<Group> groups;
    std::vector<std::pair<int, int>> gpairs;
    gpairs.reserve(s.pair_a.size());
    struct Early { int sl, a, b; };
    std::vector<Early> tmp;
    for (int lv = 0; lv < NL; ++lv) {
      const int c0 = s.level_ptr[lv];
      const int t0 = s.ltask_ptr[lv], nt = s.ltask_ptr[lv + 1] - t0;
      own[lv].reserve(nt);
      for (int t = 0; t < nt; ++t) {
        Task tk{s.task_dst[t0 + t], s.task_pos[t0 + t], (int)tpairs.size(), 0};
        if (tk.dst >= 0) {
          const int p0 = s.task_pair_ptr[t0 + t], p1 = s.task_pair_ptr[t0 + t + 1];
          const bool look = lv >= 2 && p1 - p0 > kAccMinPairs;
          tmp.clear();
          for (int p = p0; p < p1; ++p) {
            const int sl = level_of[s.blk_col[s.pair_a[p]]];  // the level that finishes both source blocks
            if (look && sl <= lv - 2) tmp.push_back(Early{sl, s.pair_a[p], s.pair_b[p]});
            else tpairs.emplace_back(s.pair_a[p], s.pair_b[p]);
          }
          if (tmp.size() > 1) std::stable_sort(tmp.begin(), tmp.end(), [](const Early &x, const Early &y) { return x.sl < y.sl; });
          for (size_t i = 0; i < tmp.size(); ++i) {
            if (i == 0 || tmp[i].sl != tmp[i - 1].sl) {
              if (i) groups.back().end = (int)gpairs.size();
              groups.push_back(Group{lv, tk.dst, tmp[i].sl, (int)gpairs.size(), 0});
            }
            gpairs.emplace_back(tmp[i].a, tmp[i].b);
          }
          if (!tmp.empty()) groups.back().end = (int)gpairs.size();
          total_pairs += p1 - p0;
        } else {
          const int j = s.level_col[c0 + tk.pos];
          for (int rr = s.row_ptr[j]; rr < s.row_ptr[j + 1]; ++rr) tpairs.emplace_back(s.row_blk[rr], -1 - s.row_col[rr]);
        }
        tk.end = (int)tpairs.size();
        load[lv] += 1 + (long long)(tk.end - tk.begin);
        own[lv].push_back(tk);
      }
    }
    // pass 2: every group goes to the latest level of its window [sl + 1, lv - 1] that still has
    // room (an even share of all products, with some slack), else to the emptiest one; the groups
    // of one block that land on the same level are one ACC task
    const long long cap = std::max<long long>(64, (total_pairs * kAccCapPct / 100) / std::max(1, NL));
    std::vector<std::pair<int, int>> placed;  // (level, group) of the block being placed
    auto flush = [&]() {
      if (placed.empty()) return;
      std::stable_sort(placed.begin(), placed.end(), [](const std::pair<int, int> &x, const std::pair<int, int> &y) { return x.first < y.first; });
      for (size_t i = 0; i < placed.size(); ++i) {
        const Group &g = groups[placed[i].second];
        if (i == 0 || placed[i].first != placed[i - 1].first) {
          if (i) early[placed[i - 1].first].back().end = (int)tpairs.size();
          early[placed[i].first].push_back(Task{g.dst, 0, (int)tpairs.size(), 0});
        }
        tpairs.insert(tpairs.end(), gpairs.begin() + g.begin, gpairs.begin() + g.end);
      }
      early[placed.back().first].back().end = (int)tpairs.size();
      placed.clear();
    };
    int cur_dst = -1;
    for (int gi = 0; gi < (int)groups.size(); ++gi) {
      const Group &g = groups[gi];
      if (g.dst != cur_dst) { flush(); cur_dst = g.dst; }
      const long long np = g.end - g.begin;
      int best = -1;
      for (int l = g.lv - 1; l > g.sl; --l)
        if (load[l] + np <= cap) { best = l; break; }
      if (best < 0) { best = g.sl + 1; for (int l = g.sl + 1; l < g.lv; ++l) if (load[l] < load[best]) best = l; }
      load[best] += np;
      placed.emplace_back(best, gi);
    }
    flush();
  }
  std::vector<std::vector<GT>> diag(C), sub(C), vec(C), acc(C), big_diag(C), big_sub(C), big_vec(C), big_acc(C);
  for (int lv = 0; lv < NL; ++lv) {
    LevelPlan &lp = plan[lv];
    const int nc = s.level_ptr[lv + 1] - s.level_ptr[lv];
    // the columns of the level are dealt over the CTAs of the cluster, heaviest first onto the
    // least loaded CTA; every task of a column runs in the column's CTA (the inverse diagonal
    // block it waits for never leaves that CTA); the ACC tasks then fill up the lightest CTAs
    std::vector<int> cta_of_pos(nc, 0);
    std::vector<long long> load(C, 0);
    if (C > 1) {
      std::vector<long long> w(nc, 0);
      for (const Task &tk : own[lv]) w[tk.pos] += 2 + (long long)(tk.end - tk.begin);
      std::vector<int> order(nc);
      std::iota(order.begin(), order.end(), 0);
      std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return w[x] > w[y]; });
      for (int pos : order) {
        const int c = (int)(std::min_element(load.begin(), load.end()) - load.begin());
        cta_of_pos[pos] = c; load[c] += w[pos];
      }
    }
    for (auto *k : {&diag, &sub, &vec, &acc, &big_diag, &big_sub, &big_vec, &big_acc}) for (auto &v : *k) v.clear();
    auto add_pairs = [&](const Task &tk) {
      GT gt{tk.dst, tk.pos, (int)lp.pairs.size(), 0};
      lp.pairs.insert(lp.pairs.end(), tpairs.begin() + tk.begin, tpairs.begin() + tk.end);
      gt.p1 = (int)lp.pairs.size();
      return gt;
    };
    for (const Task &tk : own[lv]) {
      const GT gt = add_pairs(tk);
      const bool is_diag = tk.dst >= 0 && s.blk_row[tk.dst] == s.blk_col[tk.dst];
      const int c = cta_of_pos[tk.pos];
      if (tk.dst < 0) vec[c].push_back(gt);
      else if (is_diag) (gt.p1 - gt.p0 > kSplitDiag ? big_diag : diag)[c].push_back(gt);
      else sub[c].push_back(gt);
    }
    {
      std::vector<const Task *> by_w;
      for (const Task &tk : early[lv]) by_w.push_back(&tk);
      std::stable_sort(by_w.begin(), by_w.end(), [](const Task *x, const Task *y) { return x->end - x->begin > y->end - y->begin; });
      for (const Task *tk : by_w) {
        const int c = (int)(std::min_element(load.begin(), load.end()) - load.begin());
        load[c] += 2 + (long long)(tk->end - tk->begin);
        acc[c].push_back(add_pairs(*tk));
      }
    }
    // SUB / VEC / ACC tasks with more than kSplitPairs pairs get a round of their own (pairs split
    // five ways).  Measured on B200: splitting more of them to fill idle warps makes the level SLOWER
    // (the update products are bound by the SM's shared-memory pipe, and a split round wastes lanes).
    auto npairs = [](const GT &g) { return g.p1 - g.p0; };
    auto by_size = [&](std::vector<GT> &v) { std::stable_sort(v.begin(), v.end(), [&](const GT &x, const GT &y) { return npairs(x) > npairs(y); }); };
    auto take_big = [&](std::vector<GT> &v, std::vector<GT> &big, int limit) {
      by_size(v);
      int nb = 0;
      while (nb < (int)v.size() && npairs(v[nb]) > limit) ++nb;
      big.assign(v.begin(), v.begin() + nb); v.erase(v.begin(), v.begin() + nb);
    };
    for (int c = 0; c < C; ++c) { by_size(diag[c]); take_big(sub[c], big_sub[c], kSplitPairs); take_big(vec[c], big_vec[c], kSplitPairs); take_big(acc[c], big_acc[c], kSplitAcc); }
    auto pack5 = [&](const std::vector<GT> &v, int type) {
      for (size_t i = 0; i < v.size(); i += 5) {
        Round r{type, {}};
        for (int k = 0; k < 5; ++k) r.g[k] = i + k < v.size() ? v[i + k] : GT{-1, 0, 0, 0};
        if (type == 2) for (int k = 0; k < 5; ++k) if (i + k < v.size()) r.g[k].dst = 0;  // VEC: dst >= 0 marks "active"
        lp.rounds.push_back(r);
      }
    };
    auto split5 = [&](const std::vector<GT> &v, int type) {
      for (const GT &gt : v) {
        Round r{type | 4, {}};
        const int np = gt.p1 - gt.p0;
        for (int k = 0; k < 5; ++k) {
          r.g[k] = GT{type == 2 ? 0 : gt.dst, gt.pos, gt.p0 + (int)((long long)np * k / 5), gt.p0 + (int)((long long)np * (k + 1) / 5)};
        }
        lp.rounds.push_back(r);
      }
    };
    // The DIAG rounds sit on the critical path of the level (every other round of a column waits
    // for the inverse they publish): a diagonal block with more than kSplitDiag pairs gets its
    // pairs split five ways.  SUB / VEC / ACC tasks are packed five of similar size per round (a
    // round lasts as long as its longest group), and the rounds are dealt to the warps
    // heaviest first.  Order matters for the wait-free argument: inside a CTA every DIAG round
    // precedes every SUB / VEC round (ACC rounds never wait).
    auto cost = [&](const Round &r) {
      int m = 0;
      for (int k = 0; k < 5; ++k) if (r.g[k].dst >= 0) m = std::max(m, r.g[k].p1 - r.g[k].p0);
      return ((r.type & 3) == 2 ? 1 : 3) * m + ((r.type & 4) ? 2 : 0);
    };
    for (int c = 0; c <= kSolveMaxCluster; ++c) lp.cta_rptr[c] = 0;
    for (int c = 0; c < C; ++c) {
      lp.cta_rptr[c] = (int)lp.rounds.size();
      split5(big_diag[c], 0); pack5(diag[c], 0);
      const size_t n_diag_rounds = lp.rounds.size();
      split5(big_sub[c], 1); pack5(sub[c], 1);
      split5(big_vec[c], 2); pack5(vec[c], 2);
      split5(big_acc[c], 3); pack5(acc[c], 3);
      std::stable_sort(lp.rounds.begin() + n_diag_rounds, lp.rounds.end(), [&](const Round &x, const Round &y) { return cost(x) > cost(y); });
    }
    for (int c = C; c <= kSolveMaxCluster; ++c) lp.cta_rptr[c] = (int)lp.rounds.size();
  }
  // ---- which CTAs read a block later on (as a source of an update product or of a forward
  // substitution): the CTA that finishes the block writes it into the block cache of exactly those
  std::vector<uint8_t> blk_mask(s.n_blocks, 0);
  for (int lv = 0; lv < NL; ++lv) {
    const LevelPlan &lp = plan[lv];
    for (int c = 0; c < C; ++c)
      for (int rd = lp.cta_rptr[c]; rd < lp.cta_rptr[c + 1]; ++rd) {
        const Round &r = lp.rounds[rd];
        for (int k = 0; k < 5; ++k) {
          if (r.g[k].dst < 0) continue;
          for (int p = r.g[k].p0; p < r.g[k].p1; ++p) {
            blk_mask[lp.pairs[p].first] |= (uint8_t)(1u << c);
            if ((r.type & 3) != 2) blk_mask[lp.pairs[p].second] |= (uint8_t)(1u << c);
          }
        }
      }
  }
  // ---- pass 1: segment sizes
  std::vector<int> seg_size(NL + 1, 0);
  seg_size[0] = 8 + (kSolveMaxCluster + 1) + 1;
  for (int lv = 0; lv < NL; ++lv) {
    const int c0 = s.level_ptr[lv], nc = s.level_ptr[lv + 1] - c0;
    int nb = 0;
    for (int t = 0; t < nc; ++t) { const int j = s.level_col[c0 + t]; nb += s.col_ptr[j + 1] - s.col_ptr[j] - 1; }
    const int npf = 0;  // initial values are read from global memory by the rounds themselves
    const int nbpf = lv > 0 ? level_blocks(lv - 1) : 0;
    const int nr = (int)plan[lv].rounds.size();
    seg_size[lv + 1] = 8 + (kSolveMaxCluster + 1) + 2 * nc + (nc + 1) + nb + nr + 30 * nr + 2 * (int)plan[lv].pairs.size() + 2 * npf + nbpf;
  }
  s.prog_max_seg = 0;
  for (int v : seg_size) s.prog_max_seg = std::max(s.prog_max_seg, (v + 3) & ~3);
  const SolverSmemLayout lay = solver_smem_layout(n, s.prog_max_seg);
  // ---- slot plan: block (i,k) is resident from its column's level (where it is written) until
  // the level of its row i, where it is read for the last time
  std::vector<int32_t> slot_of(s.n_blocks, -1);
  {
    std::vector<int> free_list;
    for (int v = lay.n_slots - 1; v >= 0; --v) free_list.push_back(v);
    std::vector<std::vector<int>> free_after(NL);
    for (int lv = 0; lv < NL; ++lv) {
      if (lv >= 1) for (int b : free_after[lv - 1]) free_list.push_back(slot_of[b]);
      for (int t = s.level_ptr[lv]; t < s.level_ptr[lv + 1]; ++t) {
        const int j = s.level_col[t];
        for (int b = s.col_ptr[j + 1] - 1; b > s.col_ptr[j]; --b) {  // the diagonal block is never a source
          if (free_list.empty()) break;
          slot_of[b] = free_list.back(); free_list.pop_back();
          free_after[level_of[s.blk_row[b]]].push_back(b);
        }
      }
    }
  }
  s.solver_slots = lay.n_slots;
  s.solver_cached_blocks = 0;
  for (int v : slot_of) if (v >= 0) ++s.solver_cached_blocks;
  auto ref = [&](int b) { return slot_of[b] >= 0 ? slot_of[b] : -1 - b; };
  // ---- pass 2: emit
  std::vector<int32_t> &P = s.prog;
  P.clear();
  s.prog_ptr.assign(1, 0);
  auto emit_level_blocks = [&](int lv, bool with_slots) {
    if (lv < 0 || lv >= NL) return;
    for (int t = s.level_ptr[lv]; t < s.level_ptr[lv + 1]; ++t) { const int j = s.level_col[t]; for (int b = s.col_ptr[j]; b < s.col_ptr[j + 1]; ++b) P.push_back(b); }
    if (with_slots)
      for (int t = s.level_ptr[lv]; t < s.level_ptr[lv + 1]; ++t) { const int j = s.level_col[t]; for (int b = s.col_ptr[j]; b < s.col_ptr[j + 1]; ++b) P.push_back(slot_of[b]); }
  };
  auto close_seg = [&]() { while (P.size() % 4) P.push_back(0); s.prog_ptr.push_back((int32_t)P.size()); };
  {  // prologue
    P.insert(P.end(), {0, 0, 0, 0, 0, 0, 0, 0});
    for (int c = 0; c <= kSolveMaxCluster; ++c) P.push_back(0);  // cta_rptr
    P.push_back(0);  // col_bptr[0]
    close_seg();
  }
  s.solver_rounds = 0;
  for (int lv = 0; lv < NL; ++lv) {
    const LevelPlan &lp = plan[lv];
    const int c0 = s.level_ptr[lv], nc = s.level_ptr[lv + 1] - c0;
    int nb = 0;
    for (int t = 0; t < nc; ++t) { const int j = s.level_col[c0 + t]; nb += s.col_ptr[j + 1] - s.col_ptr[j] - 1; }
    const int npf = 0;
    const int nbpf = lv > 0 ? level_blocks(lv - 1) : 0;
    const int nr = (int)lp.rounds.size(), np = (int)lp.pairs.size();
    s.solver_rounds += nr;
    P.insert(P.end(), {nc, nr, np, nb, npf, nbpf, 0, 0});
    for (int c = 0; c <= kSolveMaxCluster; ++c) P.push_back(lp.cta_rptr[c]);
    for (int t = 0; t < nc; ++t) P.push_back(s.level_col[c0 + t]);
    for (int t = 0; t < nc; ++t) P.push_back(s.col_ptr[s.level_col[c0 + t]]);
    int acc = 0;
    for (int t = 0; t < nc; ++t) { P.push_back(acc); const int j = s.level_col[c0 + t]; acc += s.col_ptr[j + 1] - s.col_ptr[j] - 1; }
    P.push_back(acc);
    for (int t = 0; t < nc; ++t) { const int j = s.level_col[c0 + t]; for (int b = s.col_ptr[j] + 1; b < s.col_ptr[j + 1]; ++b) P.push_back(s.blk_row[b]); }
    for (const Round &r : lp.rounds) P.push_back(r.type);
    for (const Round &r : lp.rounds) for (int k = 0; k < 5; ++k) P.push_back(r.g[k].dst);
    for (const Round &r : lp.rounds) for (int k = 0; k < 5; ++k) P.push_back(((r.type & 3) != 2 && r.g[k].dst >= 0) ? slot_of[r.g[k].dst] : -1);
    for (const Round &r : lp.rounds) for (int k = 0; k < 5; ++k) P.push_back(r.g[k].pos);
    for (const Round &r : lp.rounds) for (int k = 0; k < 5; ++k) P.push_back(r.g[k].p0);
    for (const Round &r : lp.rounds) for (int k = 0; k < 5; ++k) P.push_back(r.g[k].p1);
    for (const Round &r : lp.rounds) for (int k = 0; k < 5; ++k) P.push_back(((r.type & 3) == 1 && r.g[k].dst >= 0) ? blk_mask[r.g[k].dst] : 0);
    for (auto &ab : lp.pairs) P.push_back(ref(ab.first));
    for (auto &ab : lp.pairs) P.push_back(ab.second >= 0 ? ref(ab.second) : -1 - ab.second);  // VEC: column k
    emit_level_blocks(lv - 1, false);
    close_seg();
  }
  s.n_segments = NL + 1;
}

// Nested dissection by BFS level structures: the middle level of a breadth-first sweep from a
// pseudo-peripheral vertex separates the component; sub-graphs are ordered first, separators
// last.  Turns the chain-shaped elimination tree of a sliding-window (block-banded) reduced
// system into a tree of height O(bandwidth * log n).  Plays the role of cs_amd at
// linear_solver_csparse.h:262 (any order gives the same solution up to rounding).
void nested_dissection_order(int n, const std::vector<std::vector<int>> &adj_lower, std::vector<int> &perm) {
  constexpr int kLeaf = 8;
  std::vector<std::vector<int>> nb(n);
  for (int c = 0; c < n; ++c)
    for (int r : adj_lower[c]) { nb[c].push_back(r); nb[r].push_back(c); }
  perm.clear();
  perm.reserve(n);
  std::vector<int> tag(n, 0), dist(n, -1), queue;
  int cur_tag = 0;
  struct Job { std::vector<int> nodes; };
  // explicit recursion with post-order emission: (nodes, separator-to-emit-after)
  std::function<void(std::vector<int> &)> rec = [&](std::vector<int> &nodes) {
    if ((int)nodes.size() <= kLeaf) { std::sort(nodes.begin(), nodes.end()); for (int v : nodes) perm.push_back(v); return; }
    const int my = ++cur_tag;
    for (int v : nodes) tag[v] = my;
    auto bfs = [&](int src, std::vector<int> &order) {
      order.clear();
      for (int v : nodes) dist[v] = -1;
      dist[src] = 0; order.push_back(src);
      for (size_t h = 0; h < order.size(); ++h)
        for (int w : nb[order[h]]) if (tag[w] == my && dist[w] < 0) { dist[w] = dist[order[h]] + 1; order.push_back(w); }
    };
    std::vector<int> order;
    bfs(*std::min_element(nodes.begin(), nodes.end()), order);
    if (order.size() < nodes.size()) {  // disconnected: order every component on its own
      std::vector<int> comp(order), rest;
      for (int v : nodes) if (dist[v] < 0) rest.push_back(v);
      rec(comp);
      rec(rest);
      return;
    }
    bfs(order.back(), order);  // second sweep from the far end: pseudo-peripheral start
    const int depth = dist[order.back()];
    if (depth < 2) { std::sort(nodes.begin(), nodes.end()); for (int v : nodes) perm.push_back(v); return; }
    // level whose removal balances the two sides best
    std::vector<int> cnt(depth + 1, 0);
    for (int v : nodes) ++cnt[dist[v]];
    int best = -1; long long best_cost = -1, below = cnt[0];
    for (int m = 1; m < depth; ++m) {
      const long long above = (long long)nodes.size() - below - cnt[m];
      const long long cost = std::max(below, above) + 2LL * cnt[m];
      if (best < 0 || cost < best_cost) { best = m; best_cost = cost; }
      below += cnt[m];
    }
    std::vector<int> lo, hi, sep;
    for (int v : nodes) (dist[v] < best ? lo : dist[v] > best ? hi : sep).push_back(v);
    if (lo.empty() || hi.empty() || sep.size() * 2 >= nodes.size()) {
      std::sort(nodes.begin(), nodes.end()); for (int v : nodes) perm.push_back(v); return;
    }
    rec(lo);
    rec(hi);
    std::sort(sep.begin(), sep.end());
    for (int v : sep) perm.push_back(v);
  };
  std::vector<int> all(n);
  std::iota(all.begin(), all.end(), 0);
  rec(all);
}

struct SectionTimer {  // SSBA_TIMING=1: wall time of the sections of build_structure on stderr
  bool on = std::getenv("SSBA_TIMING") != nullptr;
  std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
  void mark(const char *name) {
    if (!on) return;
    const auto n = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[ssba]   %-26s %.3f ms\n", name, std::chrono::duration<double, std::milli>(n - t).count());
    t = n;
  }
};

// Elimination order, symbolic factorisation and solver program(s) of a block-sparse SPD system over n
// 6x6 block columns (adj[c] = rows r > c with a block (r, c)), shared by the bundle adjustment (the
// Schur complement) and the pose graph (H itself): natural order vs nested dissection, then the
// subtree-per-CTA re-order and program (k_tree_solve).  The level program of k_reduced_solve is only
// built when the tree program does not fit (or SSBA_SOLVER=level asks for it).
struct SolverPlan {  // what is left for finish_solver_program (run off the critical path of the structure build)
  bool want_tree = false, have_tasks = false;
  TreeAssign ta;
  std::vector<int> perm;  // q -> original column
};

bool plan_reduced_solver(int n, const std::vector<std::vector<int>> &adj, Structure &s, SolverPlan &plan, std::string &err) {
  const auto t_plan0 = std::chrono::steady_clock::now();
  struct AddTime { Structure &s; std::chrono::steady_clock::time_point t0; ~AddTime() { s.seconds_symbolic = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); } } add_time{s, t_plan0};
  static const bool force_level = [] { const char *e = std::getenv("SSBA_SOLVER"); return e && std::string(e) == "level"; }();
  std::vector<int> perm_nat(n), perm_nd, perm_tree;
  std::iota(perm_nat.begin(), perm_nat.end(), 0);
  Factor f_nat, f_nd, f_tree;
  Factor *best = nullptr;
  std::vector<int> *best_perm = &perm_nat;
  plan = SolverPlan{};
  if (n >= 24) {
    nested_dissection_order(n, adj, perm_nd);
    if (perm_nd != perm_nat) {
      if (!symbolic_factor(n, adj, perm_nd, f_nd, err, false)) return false;
      best = &f_nd; best_perm = &perm_nd;
    }
  }
  // the natural order is only worth a symbolic factorisation of its own when nested dissection
  // did not flatten the elimination tree (it always does for the block-banded windows)
  if (!best || 2 * f_nd.n_levels > n) {
    if (!symbolic_factor(n, adj, perm_nat, f_nat, err, true)) return false;
    if (best && !symbolic_factor(n, adj, perm_nd, f_nd, err, true)) return false;
    if (!best || f_nat.est_cycles <= f_nd.est_cycles) { best = &f_nat; best_perm = &perm_nat; }
    plan.have_tasks = true;
  }
  // subtree-per-CTA solver: the columns of a CTA made contiguous (a topological re-order of the same
  // elimination tree: same fill).  Candidates: the cluster size the problem asks for; a system whose factor
  // does not fit the shared memory of 8 SMs gets the 16-CTA cluster (when the device has it), if need be
  // with CTA 0 taking the top part only.  The program itself is built by finish_solver_program.
  s.tree = TreeProgram{};
  if (force_level) s.tree.why_not = "SSBA_SOLVER=level";
  if (!force_level && n > 0) {
    struct Cand { int C; bool cta0_subtree; };
    std::vector<Cand> cands = {{std::min(solver_cluster_size(n), tree_cluster_cap()), true}};
    if (n >= 64 && tree_cluster_cap() >= 16) { cands.push_back({16, true}); cands.push_back({16, false}); }
    else if (n >= 64) cands.push_back({tree_cluster_cap(), false});
    for (const Cand &cd : cands) {
      tree_assign(n, best->col_ptr, best->blk_row, cd.C, cd.cta0_subtree, plan.ta);
      if (tree_smem_estimate(n, best->col_ptr, best->blk_row, plan.ta) <= kTreeMaxSmem) { plan.want_tree = true; break; }
    }
    if (!plan.want_tree) s.tree.why_not = "estimated shared memory exceeds the cluster";
    if (plan.want_tree) {
      perm_tree.resize(n);
      for (int q = 0; q < n; ++q) perm_tree[q] = (*best_perm)[plan.ta.order[q]];
      if (!symbolic_factor(n, adj, perm_tree, f_tree, err, false)) return false;
      if (f_tree.n_blocks != best->n_blocks) { err = "internal: re-ordered factor has different fill"; return false; }
      best = &f_tree; best_perm = &perm_tree; plan.have_tasks = false;
    }
  }
  plan.perm = *best_perm;
  s.n_schur_blocks = best->n_schur;
  s.n_blocks = best->n_blocks; s.n_levels = best->n_levels; s.n_tasks = (int)best->task_dst.size();
  s.est_solver_cycles = best->est_cycles;
  s.col_ptr.swap(best->col_ptr); s.blk_row.swap(best->blk_row); s.blk_col.swap(best->blk_col);
  s.row_ptr.swap(best->row_ptr); s.row_blk.swap(best->row_blk); s.row_col.swap(best->row_col);
  s.level_ptr.swap(best->level_ptr); s.level_col.swap(best->level_col);
  s.ltask_ptr.swap(best->ltask_ptr); s.task_dst.swap(best->task_dst); s.task_pos.swap(best->task_pos); s.task_pair_ptr.swap(best->task_pair_ptr);
  s.pair_a.swap(best->pair_a); s.pair_b.swap(best->pair_b);
  return true;
}

// The solver program: the subtree-per-CTA program of k_tree_solve, else (it does not fit, or SSBA_SOLVER=level)
// the level program of k_reduced_solve with its task lists.  Reads the factor pattern of `s`, writes only
// tree / prog / prog_ptr / task lists / solver_* / solve_cluster: build_structure runs it on a thread of its own.
void finish_solver_program(int n, const std::vector<std::vector<int>> &adj, Structure &s, SolverPlan &plan) {
  const auto t0 = std::chrono::steady_clock::now();
  struct AddTime { Structure &s; std::chrono::steady_clock::time_point t0; ~AddTime() { s.seconds_symbolic += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); } } add_time{s, t0};
  if (plan.want_tree) build_tree_program(n, s.col_ptr, s.blk_row, s.row_ptr, s.row_blk, s.row_col, plan.ta, s.tree);
  if (s.tree.ok) { s.solve_cluster = s.tree.C; return; }
  if (!plan.have_tasks) {
    Factor f;
    std::string err;
    if (!symbolic_factor(n, adj, plan.perm, f, err, true)) return;  // cannot fail: the same pattern was factored before
    s.n_tasks = (int)f.task_dst.size();
    s.est_solver_cycles = f.est_cycles;
    s.ltask_ptr.swap(f.ltask_ptr); s.task_dst.swap(f.task_dst); s.task_pos.swap(f.task_pos); s.task_pair_ptr.swap(f.task_pair_ptr);
    s.pair_a.swap(f.pair_a); s.pair_b.swap(f.pair_b);
    plan.have_tasks = true;
  }
  build_solver_program(s);
}

}  // namespace

static std::atomic<int> g_ranks_on_host{1};
void set_ranks_on_host(int n) { g_ranks_on_host.store(n > 0 ? n : 1); }

// ---- a small persistent fork-join pool for the host passes over the edges (the structure build
// is on the end-to-end path of every ssba_initialize call)
namespace {
class HostPool {
 public:
  // leaked on purpose: the detached workers wait on its condition variable until process exit
  static HostPool &get() { static HostPool *p = new HostPool; return *p; }
  int size() const { return n_; }
  // fn(t, T) for t in [0, T); the caller runs t = 0.  Calls are serialised.
  // The passes of one structure build follow each other within microseconds: a worker that has just
  // finished keeps polling for the next job for a short while before it goes to sleep on the
  // condition variable (a futex wake-up costs tens of microseconds, eight times per build), and the
  // caller polls the same way for the workers to finish.
  void run(int T, const std::function<void(int, int)> &fn) {
    if (T <= 1 || n_ <= 1) { fn(0, 1); return; }
    if (T > n_) T = n_;
    std::lock_guard<std::mutex> outer(run_mu_);
    {
      std::lock_guard<std::mutex> lk(mu_);
      fn_ = &fn; T_ = T; pending_.store(T - 1, std::memory_order_relaxed);
      gen_.fetch_add(1, std::memory_order_release);
    }
    if (sleepers_.load(std::memory_order_acquire) > 0) cv_.notify_all();
    fn(0, T);
    if (!spin_until([&] { return pending_.load(std::memory_order_acquire) == 0; })) {
      std::unique_lock<std::mutex> lk(mu_);
      done_cv_.wait(lk, [&] { return pending_.load(std::memory_order_acquire) == 0; });
    }
    fn_ = nullptr;
  }

 private:
  static constexpr int kSpinMicros = 200;
  template <class Pred> static bool spin_until(Pred done) {
    const auto t_end = std::chrono::steady_clock::now() + std::chrono::microseconds(kSpinMicros);
    for (int i = 0;; ++i) {
      if (done()) return true;
      if ((i & 63) == 63 && std::chrono::steady_clock::now() > t_end) return false;
#if defined(__x86_64__) || defined(__i386__)
      __builtin_ia32_pause();
#endif
    }
  }
  HostPool() {
    // measured on the B200 hosts (16 cores) with the polling pool, cfg3: initialize 1.69 ms with 4 threads, 1.44 with 6,
    // 1.33 with 8, 1.44 with 10, 1.58 with 12, 2.1 with 16 - the passes are short and memory-bound
    int n = std::min((int)std::thread::hardware_concurrency(), 8);
    // several ranks on one host (one process per GPU) share its cores: polling workers must not oversubscribe them
    // (one core of a rank's share stays free for the planning thread of build_structure, which runs beside the pool)
    const int ranks = std::max(1, g_ranks_on_host.load());
    if (ranks > 1) n = std::max(1, std::min(n, (int)std::thread::hardware_concurrency() / ranks - 1));
    if (const char *e = std::getenv("SSBA_HOST_THREADS")) n = std::atoi(e);
    n_ = std::max(1, std::min(n, 64));
    for (int t = 1; t < n_; ++t) std::thread([this, t] { worker(t); }).detach();
  }
  void worker(int t) {
    unsigned long seen = 0;
    for (;;) {
      if (!spin_until([&] { return gen_.load(std::memory_order_acquire) != seen; })) {
        std::unique_lock<std::mutex> lk(mu_);
        sleepers_.fetch_add(1, std::memory_order_release);
        cv_.wait(lk, [&] { return gen_.load(std::memory_order_acquire) != seen; });
        sleepers_.fetch_sub(1, std::memory_order_release);
      }
      const std::function<void(int, int)> *fn = nullptr;
      int T = 0;
      {
        std::lock_guard<std::mutex> lk(mu_);
        seen = gen_.load(std::memory_order_acquire);
        if (t < T_) { fn = fn_; T = T_; }
      }
      if (fn) {
        (*fn)(t, T);
        if (pending_.fetch_sub(1, std::memory_order_acq_rel) == 1) {
          std::lock_guard<std::mutex> lk(mu_);  // pairs with the caller's wait (no lost wake-up)
          done_cv_.notify_one();
        }
      }
    }
  }
  int n_ = 1;
  std::mutex mu_, run_mu_;
  std::condition_variable cv_, done_cv_;
  const std::function<void(int, int)> *fn_ = nullptr;
  int T_ = 0;
  std::atomic<int> pending_{0}, sleepers_{0};
  std::atomic<unsigned long> gen_{0};
};
inline void split_range(int t, int T, int N, int &b, int &e) {
  b = (int)((long long)N * t / T);
  e = (int)((long long)N * (t + 1) / T);
}
}  // namespace

namespace {
// Clears a Structure for the next build but keeps the capacity of its arrays: ssba_initialize
// runs once per optimised window and re-allocating megabytes every time costs page faults.
template <class V> void clr(V &v) { v.clear(); }
void reset_keep_capacity(Structure &s) {
  clr(s.q_of_pose); clr(s.pose_of_q); clr(s.point_active);
  clr(s.slot_vertex); clr(s.slot_free); clr(s.slot_pair_ptr); clr(s.pair_vertex); clr(s.pair_q);
  clr(s.pair_edge_ptr); clr(s.pair_slot); clr(s.lchunk_slot); clr(s.lchunk_lp_ptr); clr(s.lp_pair_ptr);
  clr(s.lp_pair); clr(s.q_part_ptr); clr(s.q_part); clr(s.e_uv); clr(s.e_cam); clr(s.e_orig); clr(s.e_info);
  clr(s.e_delta); clr(s.unit_combo_ptr); clr(s.combo_blk); clr(s.blk_prod_ptr); clr(s.blk_prod); clr(s.combo_pos); clr(s.unit_slot); clr(s.unit_n); clr(s.unit_k);
  clr(s.unit_c0); clr(s.col_ptr); clr(s.blk_row); clr(s.blk_col); clr(s.ltask_ptr); clr(s.task_dst);
  clr(s.task_pos); clr(s.task_pair_ptr); clr(s.pair_a); clr(s.pair_b); clr(s.prog); clr(s.prog_ptr);
  clr(s.row_ptr); clr(s.row_blk); clr(s.row_col); clr(s.level_ptr); clr(s.level_col);
  s.n_fp = s.n_fl_global = s.n_active_edges_global = 0;
  s.n_slots = s.n_pairs = s.n_edges = s.n_fl = 0;
  s.n_lchunks = s.n_hpp_parts = s.n_units = 0;
  s.n_blocks = s.n_schur_blocks = s.n_tasks = s.n_levels = s.n_segments = 0;
  s.prog_max_seg = s.solver_slots = s.solver_cached_blocks = s.solver_rounds = 0;
  s.est_solver_cycles = 0.0;
  s.n_edges_total = 0;
}
}  // namespace

bool build_structure(const HostGraph &g, int rank, int world, Structure &s, std::string &err,
                     const std::function<void()> *on_edges_ready, const AcrossRanks *across_ranks) {
  if (across_ranks) { rank = 0; world = 1; }  // pre-sharded input: every local landmark is a slot of this rank
  SectionTimer tm;
  reset_keep_capacity(s);
  const int NK = g.n_poses, NP = g.n_points, NE = g.n_edges;
  s.n_edges_total = NE;
  if (!g.have_cams) { err = "ssba_set_cameras was not called"; return false; }
  HostPool &pool = HostPool::get();
  // small graphs: threads cost more than they save (SSBA_HOST_PAR_MIN lowers the thresholds: the CPU tests run the
  // parallel passes on their small graphs and compare with the single-thread build)
  static const int par_min = [] { const char *e = std::getenv("SSBA_HOST_PAR_MIN"); return e ? std::max(1, std::atoi(e)) : 0; }();
  const int kParEdges = par_min ? par_min : 20000, kParItems = par_min ? par_min : 8192;
  const int T = NE >= kParEdges ? pool.size() : 1;
  const int32_t *__restrict__ ge_pose = g.e_pose.data();
  const int32_t *__restrict__ ge_point = g.e_point.data();
  const uint8_t *__restrict__ ge_cam = g.e_cam.data();
  const uint8_t *__restrict__ pfix = g.pose_fixed.data();
  const uint8_t *__restrict__ lfix = g.point_fixed.data();

  // ---- validation + active sets (sparse_optimizer.cpp:201-272): an edge is active unless both
  // ends are fixed; a vertex is active when it has an active edge
  std::vector<uint8_t> pose_active(NK, 0), point_active(NP, 0);
  uvec<uint8_t> edge_active(NE);
  std::vector<int32_t> point_deg(NP + 1, 0);
  int n_active = 0;
  bool edges_by_landmark = true;  // caller's edges already grouped by non-decreasing landmark row
  {
    std::vector<int> t_active(T, 0), t_bad(T, 0), t_sorted(T, 1);
    uint8_t *pa = pose_active.data(), *la = point_active.data(), *ea = edge_active.data();
    int32_t *deg = point_deg.data();
    const int n_cams = g.cams.n;
    pool.run(T, [&](int t, int TT) {
      int e0, e1; split_range(t, TT, NE, e0, e1);
      int na = 0, prev = e0 > 0 ? ge_point[e0 - 1] : -1;
      bool sorted = true;
      // The flag arrays are shared by all threads: a flag is only written while it is still 0 (a
      // store per edge would bounce the few cache lines of pose_active between the cores), and
      // the degree of a landmark is added once per run of consecutive edges of that landmark.
      int run_l = -1, run_n = 0;
      for (int e = e0; e < e1; ++e) {
        const int ip = ge_pose[e], il = ge_point[e];
        if ((unsigned)ip >= (unsigned)NK) { t_bad[t] = 1; return; }
        if ((unsigned)il >= (unsigned)NP) { t_bad[t] = 2; return; }
        if (ge_cam[e] >= n_cams) { t_bad[t] = 3; return; }
        sorted &= il >= prev; prev = il;
        if (pfix[ip] && lfix[il]) { ea[e] = 0; continue; }
        ea[e] = 1; ++na;
        if (!__atomic_load_n(pa + ip, __ATOMIC_RELAXED)) __atomic_store_n(pa + ip, (uint8_t)1, __ATOMIC_RELAXED);
        if (il != run_l) {
          if (run_n) __atomic_fetch_add(deg + run_l, run_n, __ATOMIC_RELAXED);
          run_l = il; run_n = 0;
          if (!__atomic_load_n(la + il, __ATOMIC_RELAXED)) __atomic_store_n(la + il, (uint8_t)1, __ATOMIC_RELAXED);
        }
        ++run_n;
      }
      if (run_n) __atomic_fetch_add(deg + run_l, run_n, __ATOMIC_RELAXED);
      t_active[t] = na; t_sorted[t] = sorted;
    });
    static const char *const kBad[] = {"", "edge pose index out of range", "edge point index out of range", "edge camera index out of range"};
    for (int t = 0; t < T; ++t) if (t_bad[t]) { err = kBad[t_bad[t]]; return false; }
    for (int t = 0; t < T; ++t) { n_active += t_active[t]; edges_by_landmark = edges_by_landmark && t_sorted[t]; }
  }
  s.n_active_edges_global = n_active;
  s.n_inactive_edges_global = NE - n_active;
  if (across_ranks) {
    // which poses / landmarks are active anywhere, and the global counts
    std::vector<uint8_t> flags((size_t)NK + NP);
    std::copy(pose_active.begin(), pose_active.end(), flags.begin());
    std::copy(point_active.begin(), point_active.end(), flags.begin() + NK);
    long long sums[2] = {n_active, NE - n_active};
    if (!(*across_ranks)(flags.data(), flags.size(), sums, 2)) { err = "pre-sharded input: exchange of the active sets failed"; return false; }
    std::copy(flags.begin(), flags.begin() + NK, pose_active.begin());
    s.n_active_edges_global = (int)sums[0];
    s.n_inactive_edges_global = sums[1];
    s.point_active.assign(flags.begin() + NK, flags.end());  // global: who reports which landmark (ssba_get_points)
  } else {
    s.point_active = point_active;
  }

  tm.mark("validate + active sets");
  // ---- index mapping (sparse_optimizer.cpp:168-192): free poses in id order, then landmarks
  std::vector<int32_t> fp_of_pose(NK, -1), free_pose_rows;
  for (int i = 0; i < NK; ++i)
    if (pose_active[i] && !pfix[i]) { fp_of_pose[i] = (int)free_pose_rows.size(); free_pose_rows.push_back(i); }
  s.n_fp = (int)free_pose_rows.size();
  const int n = s.n_fp;

  // ---- edges by landmark (CSR over point rows, stable = addEdge order inside a landmark), and the number of free
  // landmarks, in one sweep over the landmarks
  std::vector<int32_t> pt_ptr(NP + 1, 0);
  {
    const uint8_t *act = across_ranks ? s.point_active.data() : point_active.data();
    int nfl = 0;
    for (int j = 0; j < NP; ++j) { pt_ptr[j + 1] = pt_ptr[j] + point_deg[j]; nfl += act[j] && !lfix[j]; }
    s.n_fl_global = nfl;
  }
  uvec<int32_t> pt_edges(n_active);
  // the order ssvio's back-end adds them in (grouped by landmark, all active): the list is the identity, written by
  // the co-visibility pass below on its way through the landmarks
  const bool identity_edges = edges_by_landmark && n_active == NE;
  if (!identity_edges) {
    std::vector<int32_t> fill(pt_ptr.begin(), pt_ptr.end() - 1);
    for (int e = 0; e < NE; ++e)
      if (edge_active[e]) pt_edges[fill[ge_point[e]]++] = e;
  }

  tm.mark("index map + edges by landmark");
  // ---- co-visibility of free poses through free landmarks = pattern of the Schur complement
  // (block_solver.hpp:224-249): per-thread bitmaps when small enough, else key lists
  const bool use_bitmap = n <= 4096;
  const size_t bm_words = use_bitmap ? ((size_t)n * n + 63) / 64 : 0;
  // The same pass collects what the landmark ORDER needs, none of which depends on the elimination order that is
  // computed from this pattern: per landmark a hash of its free-pose set (over the free-pose indices), its first
  // free pose, its number of (pose, landmark) pairs and of W pairs.  For every active landmark, because the
  // landmark order is global across ranks.
  std::vector<uint64_t> lm_hash(NP, 0);
  std::vector<int32_t> lm_minq(NP, n), lm_npairs(NP, 0), lm_k(NP, 0);  // lm_minq: first free pose (free-pose index)
  std::vector<std::vector<uint64_t>> t_bitmap(T), t_keys(T);
  pool.run(T, [&](int t, int TT) {
    int j0, j1; split_range(t, TT, NP, j0, j1);
    std::vector<uint64_t> &bm = t_bitmap[t], &keys = t_keys[t];
    if (use_bitmap) bm.assign(bm_words, 0);
    int ql[64], fl[64];
    std::vector<int32_t> big, bigf;
    auto insert_unique = [](int *lst, int &cnt, int f) {  // sorted insert, duplicates dropped
      int i = cnt;
      while (i > 0 && lst[i - 1] > f) --i;
      if (i > 0 && lst[i - 1] == f) return;
      for (int u = cnt; u > i; --u) lst[u] = lst[u - 1];
      lst[i] = f; ++cnt;
    };
    for (int j = j0; j < j1; ++j) {
      if (!point_active[j]) continue;
      const int m = pt_ptr[j + 1] - pt_ptr[j];
      int *lst = ql, *fst = fl, cnt = 0, fcnt = 0;  // distinct free poses (free-pose index) / fixed poses (row)
      if (m > 64) { big.resize(m); bigf.resize(m); lst = big.data(); fst = bigf.data(); }
      for (int k = pt_ptr[j]; k < pt_ptr[j + 1]; ++k) {
        if (identity_edges) pt_edges[k] = k;
        const int pose = ge_pose[pt_edges[k]];
        const int f = fp_of_pose[pose];
        if (f >= 0) insert_unique(lst, cnt, f); else insert_unique(fst, fcnt, pose);
      }
      lm_npairs[j] = cnt + fcnt;
      if (lfix[j]) continue;  // a fixed landmark: pairs, but no W blocks and no part in the pattern
      uint64_t h = 1469598103934665603ull;
      for (int i = 0; i < cnt; ++i) h = (h ^ (uint64_t)(lst[i] + 1)) * 1099511628211ull;
      lm_hash[j] = h; lm_k[j] = cnt;
      if (cnt) lm_minq[j] = lst[0];
      for (int a2 = 0; a2 < cnt; ++a2)
        for (int b2 = a2 + 1; b2 < cnt; ++b2) {
          const size_t bit = (size_t)lst[a2] * n + lst[b2];  // column a2 < row b2
          if (use_bitmap) bm[bit >> 6] |= 1ull << (bit & 63); else keys.push_back(bit);
        }
    }
  });
  std::vector<std::vector<int>> adj(n);  // strictly-lower rows per column, unpermuted
  if (across_ranks && !use_bitmap) { err = "pre-sharded input: more than 4096 free poses"; return false; }
  if (use_bitmap) {
    std::vector<uint64_t> &bm = t_bitmap[0];
    if (bm.empty()) bm.assign(bm_words, 0);
    for (int t = 1; t < (int)t_bitmap.size(); ++t)
      if (!t_bitmap[t].empty()) for (size_t w = 0; w < bm_words; ++w) bm[w] |= t_bitmap[t][w];
    if (across_ranks && bm_words) {
      // the pattern of the reduced system is the union over the ranks: one byte per possible block (the callback
      // takes element-wise maxima, which is OR for 0 / 1 bytes but not for packed bits)
      std::vector<uint8_t> bytes((size_t)n * n);
      for (size_t bit = 0; bit < bytes.size(); ++bit) bytes[bit] = (uint8_t)(bm[bit >> 6] >> (bit & 63) & 1);
      if (!(*across_ranks)(bytes.data(), bytes.size(), nullptr, 0)) { err = "pre-sharded input: exchange of the co-visibility pattern failed"; return false; }
      std::fill(bm.begin(), bm.end(), 0);
      for (size_t bit = 0; bit < bytes.size(); ++bit) if (bytes[bit]) bm[bit >> 6] |= 1ull << (bit & 63);
    }
    for (int c = 0; c < n; ++c)
      for (int r = c + 1; r < n; ++r) {
        const size_t bit = (size_t)c * n + r;
        if (bm[bit >> 6] >> (bit & 63) & 1) adj[c].push_back(r);
      }
  } else {
    std::vector<uint64_t> keys;
    for (auto &k : t_keys) keys.insert(keys.end(), k.begin(), k.end());
    std::sort(keys.begin(), keys.end());
    keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
    for (uint64_t k : keys) adj[(int)(k / n)].push_back((int)(k % n));
  }

  tm.mark("co-visibility");
  // ---- elimination order and symbolic factorisation over q, then the solver program: all serial and all a
  // function of the pattern only - on a thread of its own, while this one goes on with the landmark order (which
  // does not depend on q) on the pool.  `plan_state` is raised when the order and the factor pattern are in place
  // (the pair / unit lists below need them), the program follows.
  SolverPlan solver_plan;
  std::atomic<int> plan_state{0};  // 1: order + symbolic done, -1: failed (plan_err)
  std::string plan_err;
  std::thread program_thread([&] {
    const bool ok = plan_reduced_solver(n, adj, s, solver_plan, plan_err);
    if (ok) {
      const std::vector<int> &perm = solver_plan.perm;  // q -> free pose index
      s.q_of_pose.assign(NK, -1);
      s.pose_of_q.resize(n);
      for (int q = 0; q < n; ++q) { s.q_of_pose[free_pose_rows[perm[q]]] = q; s.pose_of_q[q] = free_pose_rows[perm[q]]; }
    }
    plan_state.store(ok ? 1 : -1, std::memory_order_release);
    if (ok) finish_solver_program(n, adj, s, solver_plan);
  });
  struct Joiner { std::thread &t; ~Joiner() { if (t.joinable()) t.join(); } } joiner{program_thread};
  auto find_block = [&](int row, int col) -> int {  // row >= col
    const int *b0 = s.blk_row.data() + s.col_ptr[col], *b1 = s.blk_row.data() + s.col_ptr[col + 1];
    const int *it = std::lower_bound(b0, b1, row);
    return (it != b1 && *it == row) ? (int)(it - s.blk_row.data()) : -1;
  };

  // ---- landmark order: landmarks seen by the same set of free poses are made adjacent (bucket
  // by first pose, then by the hash of the pose list), so that runs of landmarks accumulate into
  // the same Schur blocks; the shard of this rank = a contiguous range of that order, balanced
  // by edge count
  std::vector<int32_t> slots;
  {
    // counting sort by first pose, stable in the landmark index: per-thread histograms over contiguous landmark
    // ranges, offsets in (bucket, thread) order, parallel fill - the same array a serial pass would produce
    std::vector<int32_t> bucket_cnt(n + 2, 0);
    int n_act_pts = 0;
    const int TB = NP >= kParItems ? T : 1;
    std::vector<std::vector<int32_t>> t_hist(TB);
    pool.run(TB, [&](int t, int TT) {
      int j0, j1; split_range(t, TT, NP, j0, j1);
      std::vector<int32_t> &hst = t_hist[t];
      hst.assign(n + 1, 0);
      for (int j = j0; j < j1; ++j) if (point_active[j]) ++hst[lm_minq[j]];
    });
    for (int q = 0; q <= n; ++q) {
      bucket_cnt[q] = n_act_pts;
      for (int t = 0; t < TB; ++t) { if (t_hist[t].empty()) continue; const int c = t_hist[t][q]; t_hist[t][q] = n_act_pts; n_act_pts += c; }
    }
    bucket_cnt[n + 1] = n_act_pts;
    // (hash, landmark) keys side by side: the comparisons of the sort stay inside the array
    std::vector<std::pair<uint64_t, int32_t>> sorted(n_act_pts);
    pool.run(TB, [&](int t, int TT) {
      int j0, j1; split_range(t, TT, NP, j0, j1);
      std::vector<int32_t> &fill = t_hist[t];  // now: where this thread's next landmark of bucket q goes
      for (int j = j0; j < j1; ++j) if (point_active[j]) sorted[fill[lm_minq[j]]++] = {lm_hash[j], j};
    });
    // Inside a bucket the landmarks only have to be GROUPED by pose list (they arrive in
    // landmark order, which every group keeps): a stable counting sort over the few distinct
    // hashes, in order of first appearance; a comparison sort only if a bucket has many lists.
    pool.run(T, [&](int t, int TT) {
      constexpr int kMaxGroups = 32;
      std::vector<uint64_t> gh;
      std::vector<int> gid;
      std::vector<std::pair<uint64_t, int32_t>> tmp;
      for (int q = t; q <= n; q += TT) {
        std::pair<uint64_t, int32_t> *b = sorted.data() + bucket_cnt[q];
        const int m = bucket_cnt[q + 1] - bucket_cnt[q];
        if (m < 2) continue;
        gh.clear(); gid.resize(m);
        int gcnt[kMaxGroups + 1] = {0};
        bool many = false;
        for (int i = 0; i < m && !many; ++i) {
          int gi = 0;
          while (gi < (int)gh.size() && gh[gi] != b[i].first) ++gi;
          if (gi == (int)gh.size()) { if (gi == kMaxGroups) { many = true; break; } gh.push_back(b[i].first); }
          gid[i] = gi; ++gcnt[gi];
        }
        if (many) { std::sort(b, b + m); continue; }
        if (gh.size() == 1) continue;
        int off[kMaxGroups + 1];
        off[0] = 0;
        for (int gi = 0; gi < (int)gh.size(); ++gi) off[gi + 1] = off[gi] + gcnt[gi];
        tmp.assign(b, b + m);
        for (int i = 0; i < m; ++i) b[off[gid[i]]++] = tmp[i];
      }
    });
    const long long total = n_active;
    long long seen = 0;
    if (world == 1) {  // one rank owns everything: the order itself
      slots.resize(n_act_pts);
      pool.run(TB, [&](int t, int TT) {
        int i0, i1; split_range(t, TT, n_act_pts, i0, i1);
        for (int i = i0; i < i1; ++i) slots[i] = sorted[i].second;
      });
    } else
    for (const auto &hj : sorted) {
      const int j = hj.second;
      // owner = the rank whose edge-quantile holds the first edge of this landmark
      const int owner = total > 0 ? (int)std::min<long long>(world - 1, seen * world / total) : 0;
      if (owner == rank) slots.push_back(j);
      seen += point_deg[j];
    }
  }
  tm.mark("landmark order + shard");
  s.n_slots = (int)slots.size();
  s.slot_vertex.assign(slots.begin(), slots.end());
  s.slot_free.resize(s.n_slots);
  s.slot_pair_ptr.assign(s.n_slots + 1, 0);
  const bool host_values = !g.values_on_device;
  const bool have_info = host_values && !g.e_info.empty(), have_delta = host_values && !g.e_delta.empty();
  // offsets of every slot's edges, pairs and Schur targets (prefix sums), then a parallel fill
  std::vector<int32_t> slot_edge_ptr(s.n_slots + 1, 0);
  {
    // two parallel passes: the counts of every slot (random reads by landmark) and the totals of every thread's
    // range, then the prefix sums of the range on top of the totals before it
    const int TS = s.n_slots >= kParItems ? T : 1;
    std::vector<long long> t_edges(TS + 1, 0), t_pairs(TS + 1, 0), t_free(TS + 1, 0);
    int32_t *sep = slot_edge_ptr.data(), *spp = s.slot_pair_ptr.data();
    pool.run(TS, [&](int t, int TT) {
      int a, b; split_range(t, TT, s.n_slots, a, b);
      long long ne = 0, npr = 0, nf = 0;
      for (int sl = a; sl < b; ++sl) {
        const int j = slots[sl];
        sep[sl + 1] = point_deg[j]; spp[sl + 1] = lm_npairs[j];
        ne += point_deg[j]; npr += lm_npairs[j];
        s.slot_free[sl] = !lfix[j];
        nf += !lfix[j];
      }
      t_edges[t + 1] = ne; t_pairs[t + 1] = npr; t_free[t + 1] = nf;
    });
    for (int t = 0; t < TS; ++t) { t_edges[t + 1] += t_edges[t]; t_pairs[t + 1] += t_pairs[t]; s.n_fl += (int)t_free[t + 1]; }
    pool.run(TS, [&](int t, int TT) {
      int a, b; split_range(t, TT, s.n_slots, a, b);
      int32_t ce = (int32_t)t_edges[t], cp = (int32_t)t_pairs[t];
      for (int sl = a; sl < b; ++sl) { ce += sep[sl + 1]; cp += spp[sl + 1]; sep[sl + 1] = ce; spp[sl + 1] = cp; }
    });
  }
  tm.mark("slot offsets");
  // ---- from here on the elimination order is needed (pairs are sorted by q inside a landmark, the Schur targets are
  // blocks of the factor pattern): wait for the planning thread - normally long done
  for (int spin = 0; plan_state.load(std::memory_order_acquire) == 0; ++spin) {
    if (spin < 4096) {
#if defined(__x86_64__) || defined(__i386__)
      __builtin_ia32_pause();
#endif
    } else {
      std::this_thread::yield();
    }
  }
  if (plan_state.load(std::memory_order_acquire) < 0) { err = plan_err; return false; }
  const int32_t *__restrict__ qmap = s.q_of_pose.data();
  tm.mark("wait for order + symbolic");
  {
    const size_t ne_local = (size_t)slot_edge_ptr[s.n_slots], np_local = (size_t)s.slot_pair_ptr[s.n_slots];
    s.n_edges = (int)ne_local; s.n_pairs = (int)np_local;
    s.e_orig.resize(ne_local); s.e_uv.resize(host_values ? 2 * ne_local : 0); s.e_cam.resize(ne_local);
    if (have_info) s.e_info.resize(3 * ne_local);
    if (have_delta) s.e_delta.resize(ne_local);
    s.pair_vertex.resize(np_local); s.pair_q.resize(np_local); s.pair_edge_ptr.resize(np_local + 1);
    s.pair_slot.resize(np_local);
    s.pair_edge_ptr[np_local] = (int32_t)ne_local;
    pool.run(T, [&](int t, int TT) {
      int s0, s1; split_range(t, TT, s.n_slots, s0, s1);
      const double *__restrict__ ge_uv = g.e_uv.data();
      int32_t *__restrict__ pe = pt_edges.data();
      int32_t *__restrict__ o_orig = s.e_orig.data();
      std::vector<std::pair<long long, int32_t>> big;
      double *__restrict__ o_uv = s.e_uv.data();
      uint8_t *__restrict__ o_cam = s.e_cam.data();
      int32_t *__restrict__ o_pv = s.pair_vertex.data(), *__restrict__ o_pq = s.pair_q.data(),
              *__restrict__ o_pe = s.pair_edge_ptr.data(),
              *__restrict__ o_ps = s.pair_slot.data();
      for (int sl = s0; sl < s1; ++sl) {
        const int j = slots[sl];
        if (sl + 6 < s1) {  // the class order visits landmarks far apart in the caller's arrays
          const int e_next = pe[pt_ptr[slots[sl + 6]]];
          if (host_values) {
            __builtin_prefetch(ge_uv + 2 * (size_t)e_next);
            __builtin_prefetch(ge_uv + 2 * (size_t)e_next + 8);
            __builtin_prefetch(ge_uv + 2 * (size_t)e_next + 16);
          }
          __builtin_prefetch(ge_pose + e_next);
          __builtin_prefetch(ge_cam + e_next);
        }
        if (sl + 12 < s1) __builtin_prefetch(pe + pt_ptr[slots[sl + 12]]);
        size_t ne = (size_t)slot_edge_ptr[sl], npair = (size_t)s.slot_pair_ptr[sl];
        int last_pose = -1;
        const int k1 = pt_ptr[j + 1];
        {
          // the landmark's edges sorted by pose, in place: free poses first by q, then fixed poses by row; addEdge
          // order inside a pose (stable)
          int32_t *seg = pe + pt_ptr[j];
          const int m = k1 - pt_ptr[j];
          auto pose_key = [&](int e) -> long long { const int q = qmap[ge_pose[e]]; return q >= 0 ? q : (long long)n + ge_pose[e]; };
          if (m <= 48) {  // insertion sort, stable
            long long keys[48];
            for (int i = 0; i < m; ++i) {
              const int e = seg[i];
              const long long k = pose_key(e);
              int u = i;
              while (u > 0 && keys[u - 1] > k) { keys[u] = keys[u - 1]; seg[u] = seg[u - 1]; --u; }
              keys[u] = k; seg[u] = e;
            }
          } else {
            big.clear();
            for (int i = 0; i < m; ++i) big.emplace_back(pose_key(seg[i]), seg[i]);
            std::sort(big.begin(), big.end());
            for (int i = 0; i < m; ++i) seg[i] = big[i].second;
          }
        }
        for (int k = pt_ptr[j]; k < k1; ++k) {
          const int e = pe[k];
          const int pose = ge_pose[e];
          if (pose != last_pose) {  // edges are sorted by pose: a new (pose, landmark) pair starts
            last_pose = pose;
            const int q = qmap[pose];
            o_pv[npair] = pose; o_pq[npair] = q; o_pe[npair] = (int32_t)ne; o_ps[npair] = sl; ++npair;
          }
          o_orig[ne] = e;
          if (host_values) { o_uv[2 * ne] = ge_uv[2 * (size_t)e]; o_uv[2 * ne + 1] = ge_uv[2 * (size_t)e + 1]; }
          o_cam[ne] = ge_cam[e];
          if (have_info) for (int u = 0; u < 3; ++u) s.e_info[3 * ne + u] = g.e_info[3 * (size_t)e + u];
          if (have_delta) s.e_delta[ne] = g.e_delta[e];
          ++ne;
        }
      }
    });
  }
  tm.mark("slots / pairs / edges fill");
  // the per-edge / per-pair arrays are final: the caller may start uploading them while the
  // solver program and the small index lists are still being built
  if (on_edges_ready) (*on_edges_ready)();
  // ---- CTAs of the per-pair kernels: runs of whole landmarks with <= kLinPairs pairs
  {
    s.lchunk_slot.assign(1, 0);
    int acc = 0;
    for (int sl = 0; sl < s.n_slots; ++sl) {
      const int np_ = s.slot_pair_ptr[sl + 1] - s.slot_pair_ptr[sl];
      if (acc > 0 && acc + np_ > kLinPairs) { s.lchunk_slot.push_back(sl); acc = 0; }
      acc += np_;
    }
    if (s.n_slots > 0) s.lchunk_slot.push_back(s.n_slots);
    s.n_lchunks = (int)s.lchunk_slot.size() - 1;
  }
  tm.mark("  chunks");

  // ---- Schur work units: a run of <= kSchurRun consecutive free landmarks with the same W pose list
  // x a chunk of <= 32 of its k(k+1)/2 block pairs (one lane per block pair, one warp per unit)
  {
    auto wcount = [&](int sl) { return lm_k[slots[sl]]; };
    auto same_list = [&](int x, int y, int k) {
      if (lm_hash[slots[x]] != lm_hash[slots[y]]) return false;
      for (int i = 0; i < k; ++i)
        if (s.pair_q[s.slot_pair_ptr[x] + i] != s.pair_q[s.slot_pair_ptr[y] + i]) return false;
      return true;
    };
    // runs never cross the boundaries of fixed blocks of slots (independent of the thread count, so the units -
    // and with them the summation order of the reduced system - only depend on the graph): the blocks are scanned
    // in parallel and their unit lists concatenated in block order
    constexpr int kUnitBlock = 1024;
    const int n_ublocks = (s.n_slots + kUnitBlock - 1) / kUnitBlock;
    struct UnitList { std::vector<int32_t> slot, n, k, c0; };
    std::vector<UnitList> ul(n_ublocks);
    pool.run(std::min(T, std::max(1, n_ublocks)), [&](int t, int TT) {
      for (int ub = t; ub < n_ublocks; ub += TT) {
        UnitList &u = ul[ub];
        const int end = std::min(s.n_slots, (ub + 1) * kUnitBlock);
        int sl = ub * kUnitBlock;
        while (sl < end) {
          const int k = wcount(sl);
          if (k == 0) { ++sl; continue; }
          int e = sl + 1;
          // the run's W blocks are staged in shared memory by k_schur: at most kSchurRunPairs of them
          const int max_run = std::max(1, std::min(kSchurRun, kSchurRunPairs / k));
          while (e < end && e - sl < max_run && wcount(e) == k && same_list(sl, e, k)) ++e;
          const int npairs = k * (k + 1) / 2;
          for (int c0 = 0; c0 < npairs; c0 += 32) { u.slot.push_back(sl); u.n.push_back(e - sl); u.k.push_back(k); u.c0.push_back(c0); }
          sl = e;
        }
      }
    });
    for (const UnitList &u : ul) {
      s.unit_slot.insert(s.unit_slot.end(), u.slot.begin(), u.slot.end());
      s.unit_n.insert(s.unit_n.end(), u.n.begin(), u.n.end());
      s.unit_k.insert(s.unit_k.end(), u.k.begin(), u.k.end());
      s.unit_c0.insert(s.unit_c0.end(), u.c0.begin(), u.c0.end());
    }
    s.n_units = (int)s.unit_slot.size();
    tm.mark("  unit runs");
    // Schur targets of every unit: for its block pairs a <= b (poses of the run sorted by q):
    // block (row q_b, col q_a) of the factor pattern
    s.unit_combo_ptr.assign(s.n_units + 1, 0);
    for (int u = 0; u < s.n_units; ++u) {
      const int k = s.unit_k[u];
      s.unit_combo_ptr[u + 1] = s.unit_combo_ptr[u] + std::min(32, k * (k + 1) / 2 - s.unit_c0[u]);
    }
    s.combo_blk.resize((size_t)s.unit_combo_ptr[s.n_units]);
    std::vector<int> t_bad(T, 0);
    pool.run(T, [&](int t, int TT) {
      int u0, u1; split_range(t, TT, s.n_units, u0, u1);
      for (int u = u0; u < u1; ++u) {
        const int k = s.unit_k[u], c0 = s.unit_c0[u], c1 = c0 + (s.unit_combo_ptr[u + 1] - s.unit_combo_ptr[u]);
        const int32_t *qs = s.pair_q.data() + s.slot_pair_ptr[s.unit_slot[u]];
        int idx = 0, o = s.unit_combo_ptr[u];
        for (int a2 = 0; a2 < k && idx < c1; ++a2)
          for (int b2 = a2; b2 < k && idx < c1; ++b2, ++idx) {
            if (idx < c0) continue;
            const int blk = find_block(qs[b2], qs[a2]);
            if (blk < 0) { t_bad[t] = 1; return; }
            s.combo_blk[o++] = blk;
          }
      }
    });
    for (int t = 0; t < T; ++t) if (t_bad[t]) { err = "internal: Schur block missing from the factor pattern"; return false; }
    tm.mark("  combos");
    // producers of every factor block: a stable counting sort of the combos by block
    const int ncomb = (int)s.combo_blk.size();
    s.blk_prod_ptr.assign(s.n_blocks + 1, 0);
    for (int c = 0; c < ncomb; ++c) ++s.blk_prod_ptr[s.combo_blk[c] + 1];
    for (int b = 0; b < s.n_blocks; ++b) s.blk_prod_ptr[b + 1] += s.blk_prod_ptr[b];
    s.blk_prod.resize(ncomb);
    {
      std::vector<int32_t> fill(s.blk_prod_ptr.begin(), s.blk_prod_ptr.end() - 1);
      for (int c = 0; c < ncomb; ++c) s.blk_prod[fill[s.combo_blk[c]]++] = c;
    }
    s.combo_pos.resize(ncomb);
    for (int p = 0; p < ncomb; ++p) s.combo_pos[s.blk_prod[p]] = p;
  }

  tm.mark("chunks + schur units");
  // ---- Hpp partials of the linearize CTAs: per chunk its distinct free poses ("local poses")
  // with the list of the chunk's pairs on each, and per pose the list of its partials.
  // Two parallel passes over the chunks (count, then fill at prefix offsets).
  {
    const int NC = s.n_lchunks;
    std::vector<int32_t> nlp(NC + 1, 0), nlist(NC + 1, 0);
    auto chunk_pass = [&](int t, int TT, bool fill, std::vector<int32_t> *lp_q) {
      int c0, c1; split_range(t, TT, NC, c0, c1);
      std::vector<int32_t> stamp(n, -1), local(n, 0), first_seen;
      std::vector<std::vector<uint8_t>> lists;
      for (int c = c0; c < c1; ++c) {
        const int a0 = s.slot_pair_ptr[s.lchunk_slot[c]], a1 = s.slot_pair_ptr[s.lchunk_slot[c + 1]];
        const bool small = a1 - a0 <= kLinPairs;
        first_seen.clear();
        if (fill) { for (auto &l : lists) l.clear(); }
        int nl = 0, npairs_free = 0;
        for (int a = a0; a < a1; ++a) {
          const int q = s.pair_q[a];
          if (q < 0) continue;
          ++npairs_free;
          if (!small) { if (fill) (*lp_q)[nlp[c] + nl] = q; ++nl; continue; }  // big landmark: one partial per pair
          if (stamp[q] != c) {
            stamp[q] = c; local[q] = nl++;
            if (fill) { if ((int)lists.size() < nl) lists.emplace_back(); (*lp_q)[nlp[c] + local[q]] = q; }
          }
          if (fill) lists[local[q]].push_back((uint8_t)(a - a0));
        }
        if (!fill) { nlp[c + 1] = nl; nlist[c + 1] = small ? npairs_free : 0; }
        else {
          int off = nlist[c];
          for (int i = 0; i < nl; ++i) {
            if (small) { std::copy(lists[i].begin(), lists[i].end(), s.lp_pair.begin() + off); off += (int)lists[i].size(); }
            s.lp_pair_ptr[nlp[c] + i + 1] = off;
          }
        }
        // stamps are keyed by chunk id, so they need no reset between chunks
      }
    };
    pool.run(T, [&](int t, int TT) { chunk_pass(t, TT, false, nullptr); });
    for (int c = 0; c < NC; ++c) { nlp[c + 1] += nlp[c]; nlist[c + 1] += nlist[c]; }
    s.n_hpp_parts = nlp[NC];
    s.lchunk_lp_ptr.assign(nlp.begin(), nlp.end());
    s.lp_pair.resize(nlist[NC]);
    s.lp_pair_ptr.assign(s.n_hpp_parts + 1, 0);
    std::vector<int32_t> lp_q(s.n_hpp_parts);
    pool.run(T, [&](int t, int TT) { chunk_pass(t, TT, true, &lp_q); });
    s.q_part_ptr.assign(n + 1, 0);
    for (int q : lp_q) ++s.q_part_ptr[q + 1];
    for (int q = 0; q < n; ++q) s.q_part_ptr[q + 1] += s.q_part_ptr[q];
    s.q_part.resize(lp_q.size());
    std::vector<int32_t> fill(s.q_part_ptr.begin(), s.q_part_ptr.end() - 1);
    for (size_t i = 0; i < lp_q.size(); ++i) s.q_part[fill[lp_q[i]]++] = (int32_t)i;
  }
  tm.mark("hpp partial lists");
  program_thread.join();
  tm.mark("solver program (rest)");
  return true;
}

// The reduced-solver part of a Structure for an arbitrary block-sparse SPD system over `n` 6x6 block
// columns (pose-graph optimisation: H itself instead of a Schur complement): elimination order,
// symbolic factorisation, levels, the packed device program.  adj[c] = rows r > c with a block (r, c).
// perm_out[q] = original column of elimination position q.
bool build_solver_structure(int n, const std::vector<std::vector<int>> &adj, Structure &s, std::vector<int> &perm_out,
                            std::string &err) {
  reset_keep_capacity(s);
  s.n_fp = n;
  SolverPlan plan;
  if (!plan_reduced_solver(n, adj, s, plan, err)) return false;
  finish_solver_program(n, adj, s, plan);
  perm_out = plan.perm;
  return true;
}

void parallel_copy(const std::vector<CopyJob> &jobs) {
  size_t total = 0;
  for (auto &j : jobs) total += j.bytes;
  HostPool &pool = HostPool::get();
  const int T = total >= (1u << 20) ? pool.size() : 1;
  // cut every job into 256 KiB pieces, deal the pieces round-robin
  struct Piece { char *d; const char *s; size_t n; };
  std::vector<Piece> pieces;
  for (auto &j : jobs)
    for (size_t o = 0; o < j.bytes; o += (256u << 10))
      pieces.push_back({(char *)j.dst + o, (const char *)j.src + o, std::min<size_t>(256u << 10, j.bytes - o)});
  pool.run(T, [&](int t, int TT) {
    for (size_t i = t; i < pieces.size(); i += TT) std::memcpy(pieces[i].d, pieces[i].s, pieces[i].n);
  });
}

bool parallel_equal(const std::vector<CopyJob> &jobs) {
  size_t total = 0;
  for (auto &j : jobs) total += j.bytes;
  HostPool &pool = HostPool::get();
  const int T = total >= (1u << 20) ? pool.size() : 1;
  struct Piece { const char *a, *b; size_t n; };
  std::vector<Piece> pieces;
  for (auto &j : jobs)
    for (size_t o = 0; o < j.bytes; o += (256u << 10))
      pieces.push_back({(const char *)j.dst + o, (const char *)j.src + o, std::min<size_t>(256u << 10, j.bytes - o)});
  std::atomic<int> differ{0};
  pool.run(T, [&](int t, int TT) {
    for (size_t i = t; i < pieces.size() && !differ.load(std::memory_order_relaxed); i += TT)
      if (std::memcmp(pieces[i].a, pieces[i].b, pieces[i].n) != 0) differ.store(1, std::memory_order_relaxed);
  });
  return differ.load() == 0;
}

// dst[k * i .. k * i + k) = src[k * idx[i] ..] for i < n (k doubles per element), on the host thread pool
void parallel_gather_doubles(double *dst, const double *src, const int32_t *idx, size_t n, int k) {
  HostPool &pool = HostPool::get();
  const int T = n >= 50000 ? pool.size() : 1;
  pool.run(T, [&](int t, int TT) {
    int b, e; split_range(t, TT, (int)n, b, e);
    if (k == 2) for (int i = b; i < e; ++i) { const size_t o = 2 * (size_t)idx[i]; dst[2 * (size_t)i] = src[o]; dst[2 * (size_t)i + 1] = src[o + 1]; }
    else for (int i = b; i < e; ++i) for (int u = 0; u < k; ++u) dst[(size_t)k * i + u] = src[(size_t)k * idx[i] + u];
  });
}

bool plan_shards(const HostGraph &g, int world, std::vector<int32_t> &owner, std::string &err) {
  owner.assign(g.n_points, -1);
  for (int r = 0; r < world; ++r) {
    Structure s;
    if (!build_structure(g, r, world, s, err)) return false;
    for (int v : s.slot_vertex) owner[v] = r;
  }
  return true;
}

}  // namespace ssba
