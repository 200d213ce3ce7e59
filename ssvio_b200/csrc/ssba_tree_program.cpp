// ssba_tree_program.cpp — see ssba_tree_program.hpp.
#include "ssba_tree_program.hpp"
#include "ssba_structure.hpp"

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <numeric>

namespace ssba {

// ---- which CTA factors which column, and in which step --------------------------------------------
// Works on the factor pattern of ANY valid elimination order (children before parents) and returns a
// topological re-order in which the columns of a CTA are contiguous ([CTA 1 .. C-1 | CTA 0 | top], by
// step inside a CTA), so that after a second symbolic factorisation under that order every CTA's factor
// blocks and right-hand sides are one contiguous slice of the reduced system (one bulk copy each).
static std::atomic<int> g_tree_cluster_cap{[] {
  const char *e = std::getenv("SSBA_TREE_CLUSTER_CAP");  // tests of the host logic; the library asks the device
  const int c = e ? std::atoi(e) : 8;
  return c >= 1 && c <= kTreeMaxCluster ? c : 8;
}()};
void set_tree_cluster_cap(int cap) { g_tree_cluster_cap.store(std::max(1, std::min(cap, kTreeMaxCluster))); }
int tree_cluster_cap() { return g_tree_cluster_cap.load(); }

void tree_assign(int n, const std::vector<int32_t> &col_ptr, const std::vector<int32_t> &blk_row, int C_want, bool cta0_subtree,
                 TreeAssign &a) {
  a = TreeAssign{};
  a.C = 1;
  std::vector<int> parent(n, -1);
  std::vector<std::vector<int>> children(n);
  std::vector<long long> W(n, 0);
  for (int j = 0; j < n; ++j) {
    const int nb = col_ptr[j + 1] - col_ptr[j];
    if (nb > 1) { parent[j] = blk_row[col_ptr[j] + 1]; children[parent[j]].push_back(j); }
    W[j] += 4 + (long long)nb * (nb + 1) / 2;  // ~ block products the column causes
  }
  for (int j = 0; j < n; ++j) if (parent[j] >= 0) W[parent[j]] += W[j];  // children precede parents
  // ---- cut: move the root of the heaviest subtree into the top part until there are C subtrees
  std::vector<int> active;
  std::vector<uint8_t> top(n, 0);
  for (int j = 0; j < n; ++j) if (parent[j] < 0) active.push_back(j);
  const int n_bins_want = cta0_subtree ? C_want : C_want - 1;  // CTAs that take subtrees
  if (C_want > 1) {
    while ((int)active.size() < n_bins_want) {
      int best = -1;
      for (int i = 0; i < (int)active.size(); ++i)
        if (!children[active[i]].empty() && (best < 0 || W[active[i]] > W[active[best]])) best = i;
      if (best < 0) break;
      const int r = active[best];
      active.erase(active.begin() + best);
      top[r] = 1;
      for (int c : children[r]) active.push_back(c);
    }
  }
  // ---- subtrees onto CTAs: heaviest first onto the least loaded
  std::stable_sort(active.begin(), active.end(), [&](int x, int y) { return W[x] > W[y]; });
  const int first_bin = (cta0_subtree || C_want == 1) ? 0 : 1;
  const int C = std::max(1, std::min(C_want, (int)active.size() + first_bin));
  std::vector<long long> load(C, 0);
  std::vector<int> cta(n, 0);
  {
    std::vector<int> root_cta(n, -1);
    const int fb = C > 1 ? first_bin : 0;
    for (int r : active) {
      const int c = (int)(std::min_element(load.begin() + fb, load.end()) - load.begin());
      root_cta[r] = c; load[c] += W[r];
    }
    for (int j = n - 1; j >= 0; --j) {  // parents before children
      if (top[j]) { cta[j] = 0; continue; }
      cta[j] = root_cta[j] >= 0 ? root_cta[j] : cta[parent[j]];
    }
  }
  if (C == 1) std::fill(top.begin(), top.end(), 0);  // one CTA: no top part, no hand-off
  // ---- elimination levels inside every CTA's forest / inside the top part
  std::vector<int> lev(n, 0);
  for (int j = 0; j < n; ++j)
    for (int c : children[j]) if (top[c] == top[j]) lev[j] = std::max(lev[j], lev[c] + 1);
  // steps: a level of a CTA is cut into chunks of <= kTreeStepCols columns
  std::vector<int> step(n, 0);
  int n_steps_a0 = 0;
  for (int c = 0; c < C; ++c) {
    std::vector<int> cols;
    for (int j = 0; j < n; ++j) if (!top[j] && cta[j] == c) cols.push_back(j);
    std::stable_sort(cols.begin(), cols.end(), [&](int x, int y) { return lev[x] < lev[y]; });
    int s = -1, in_step = 0, cur_lev = -1;
    for (int j : cols) {
      if (lev[j] != cur_lev || in_step == kTreeStepCols) { ++s; in_step = 0; cur_lev = lev[j]; }
      step[j] = s; ++in_step;
    }
    a.steps_a[c] = s + 1 + 1;  // + the flush step
    if (c == 0) n_steps_a0 = a.steps_a[0];
  }
  {
    std::vector<int> cols;
    for (int j = 0; j < n; ++j) if (top[j]) cols.push_back(j);
    std::stable_sort(cols.begin(), cols.end(), [&](int x, int y) { return lev[x] < lev[y]; });
    int s = n_steps_a0 - 1, in_step = 0, cur_lev = -1;
    for (int j : cols) {
      if (lev[j] != cur_lev || in_step == kTreeStepCols) { ++s; in_step = 0; cur_lev = lev[j]; }
      step[j] = s; ++in_step;
    }
    a.steps_b = s - (n_steps_a0 - 1);
    a.n_top = (int)cols.size();
  }
  // ---- the new order
  a.C = C;
  a.order.clear();
  a.order.reserve(n);
  auto emit = [&](int c, bool t) {
    std::vector<int> cols;
    for (int j = 0; j < n; ++j) if ((top[j] != 0) == t && cta[j] == c) cols.push_back(j);
    std::stable_sort(cols.begin(), cols.end(), [&](int x, int y) { return step[x] < step[y]; });
    a.order.insert(a.order.end(), cols.begin(), cols.end());
  };
  for (int c = 1; c < C; ++c) emit(c, false);
  emit(0, false);
  emit(0, true);
  a.cta.resize(n); a.step.resize(n); a.top.resize(n);
  for (int q = 0; q < n; ++q) { const int j = a.order[q]; a.cta[q] = cta[j]; a.step[q] = step[j]; a.top[q] = top[j]; }
}

// Upper estimate of the dynamic shared memory build_tree_program will need under an assignment (pattern in
// the OLD order, `old_of` = a.order): exact for the factor blocks and vectors, bounds for the contribution
// slots and the program.  Lets the caller pick a cluster shape before the program itself is built (which
// happens off the critical path of the structure build).
size_t tree_smem_estimate(int n, const std::vector<int32_t> &col_ptr, const std::vector<int32_t> &blk_row, const TreeAssign &a) {
  const int C = a.C;
  std::vector<double> words(C, kTH_Words + 64.0);
  std::vector<long long> nblk(C, 0), ncols(C, 0), nbound(C, 0), ntemp(C, 0), in_step(C, 0);
  std::vector<int> stamp(n, -1), cur_step(C, -1);
  std::vector<int> new_of(n);
  for (int q = 0; q < n; ++q) new_of[a.order[q]] = q;
  for (int q = 0; q < n; ++q) {
    const int j = a.order[q], c = a.cta[q];
    const int m = col_ptr[j + 1] - col_ptr[j] - 1;
    nblk[c] += m + 1; ++ncols[c];
    {  // unscaled copies of the sub-diagonal blocks of one step's columns (columns of a step are adjacent in the order)
      const int key = a.step[q] + (a.top[q] ? 1 << 20 : 0);
      if (key != cur_step[c]) { cur_step[c] = key; in_step[c] = 0; }
      in_step[c] += m;
      ntemp[c] = std::max(ntemp[c], in_step[c]);
    }
    words[c] += 3.6 * (0.5 * m * (m + 1) + m) + 2.4 * (m + 1) + 12 + 3 * m;  // (backward: a pair word per block, an item per destination and source step)
    if (c)
      for (int b = col_ptr[j] + 1; b < col_ptr[j + 1]; ++b) {
        const int i = blk_row[b];
        if (a.top[new_of[i]] && stamp[i] != c) { stamp[i] = c; ++nbound[c]; }
      }
  }
  size_t worst = 0;
  for (int c = 0; c < C; ++c) {
    words[c] += 8.0 * (a.steps_a[c] + (c == 0 ? a.steps_b : 0));
    if (c) { words[0] += (double)(nbound[c] * (nbound[c] + 1) / 2 + nbound[c]); words[c] += (double)nbound[c]; }
  }
  for (int c = 0; c < C; ++c) {
    const long long contrib = c ? 36 * (nbound[c] * (nbound[c] + 1) / 2) + 6 * nbound[c] : 0;
    const size_t bytes = 8 * (size_t)(36 * nblk[c] + 12 * ncols[c] + contrib + 36 * ntemp[c]) + 4 * (size_t)words[c] + kTreeMiscBytes + 64;
    worst = std::max(worst, bytes);
  }
  return worst;
}

namespace {

struct Item { int dest, nrows, p0, p1; };   // product item: dest -= sum over pairs of A B^T (row-wise)
// product item: { dest | nrows << 16 | n_pairs << 20, first pair word }; panel item: { dest | nrows << 16, diagonal block | unscaled copy << 16 }
inline bool push_product_item(std::vector<int32_t> &w, int dest, int nrows, int p0, int np) {
  if (np >= 4096) return false;
  w.push_back((int32_t)((unsigned)dest | ((unsigned)nrows << 16) | ((unsigned)np << 20)));
  w.push_back(p0);
  return true;
}
inline void push_panel_item(std::vector<int32_t> &w, int dest, int nrows, int diag, int xcopy) {
  w.push_back((int32_t)((unsigned)dest | ((unsigned)nrows << 16)));
  w.push_back((int32_t)((unsigned)diag | ((unsigned)xcopy << 16)));
}
// one term of the backward substitution: w_dest -= Y^T x, `blk` / `x` = pool offsets; `rank` orders the destinations
struct BwdPair { int dest, rank, src, blk, x; };
// pair words, then rounds of five items { destination vector | n_pairs << 16, first pair word }; returns the rounds
inline int emit_bwd_rounds(std::vector<int32_t> &w, std::vector<BwdPair> &bp, int V0, int q0, int &off_items) {
  std::stable_sort(bp.begin(), bp.end(), [](const BwdPair &x, const BwdPair &y) {
    if (x.rank != y.rank) return x.rank < y.rank;
    if (x.dest != y.dest) return x.dest > y.dest;  // later columns become final first
    return x.src > y.src;
  });
  const int off_pairs = (int)w.size();
  for (const BwdPair &p : bp) w.push_back((int32_t)((unsigned)p.blk | ((unsigned)p.x << 16)));
  if (w.size() & 1) w.push_back(0);
  std::vector<int32_t> items;
  for (size_t i = 0; i < bp.size();) {
    size_t j = i;
    while (j < bp.size() && bp[j].dest == bp[i].dest) ++j;
    if (j - i >= 65536) return -1;
    items.push_back((int32_t)((unsigned)(V0 + 6 * (bp[i].dest - q0)) | ((unsigned)(j - i) << 16)));
    items.push_back(off_pairs + (int)i);
    i = j;
  }
  const int n_rounds = ((int)items.size() / 2 + 4) / 5;
  items.resize((size_t)n_rounds * kTreeRoundWords, 0);
  off_items = (int)w.size();  // even: the items are read as 8-byte words
  w.insert(w.end(), items.begin(), items.end());
  return n_rounds;
}

}  // namespace

bool build_tree_program(int n, const std::vector<int32_t> &col_ptr, const std::vector<int32_t> &blk_row,
                        const std::vector<int32_t> &row_ptr, const std::vector<int32_t> &row_blk,
                        const std::vector<int32_t> &row_col, const TreeAssign &a, TreeProgram &tp) {
  tp = TreeProgram{};
  const int C = a.C;
  tp.C = C;
  tp.n_top_cols = a.n_top;
  auto fail = [&](const char *why) { tp.ok = false; tp.why_not = why; return false; };
  if (C < 1 || C > kTreeMaxCluster) return fail("cluster size");
  // ---- column ranges (the order of tree_assign: CTA 1 .. C-1, CTA 0, top)
  std::vector<int> q0(C, 0), q1(C, 0);
  {
    int q = 0;
    for (int c = 1; c < C; ++c) { q0[c] = q; while (q < n && a.cta[q] == c && !a.top[q]) ++q; q1[c] = q; }
    q0[0] = q; q1[0] = n;
    for (int qq = q; qq < n; ++qq) if (a.cta[qq] != 0) return fail("internal: column order");
  }
  std::vector<int> V0(C), BP0(C), CB0(C);
  for (int c = 0; c < C; ++c) {
    tp.q0[c] = q0[c]; tp.n_own_cols[c] = q1[c] - q0[c];
    tp.b0[c] = col_ptr[q0[c]]; tp.n_own_blocks[c] = col_ptr[q1[c]] - col_ptr[q0[c]];
    V0[c] = 36 * tp.n_own_blocks[c];
    BP0[c] = V0[c] + 6 * tp.n_own_cols[c];
    CB0[c] = BP0[c] + 6 * tp.n_own_cols[c];
  }
  auto own = [&](int c, int q) { return q >= q0[c] && q < q1[c]; };
  // unscaled copies: the sub-diagonal blocks X_ik of the columns of one step sit side by side in the CTA's
  // scratch area from the panel of their step until the products of the next step have read them
  std::vector<int32_t> xcopy_idx(col_ptr[n], -1);
  std::vector<int> xcopy_blocks(C, 0);
  for (int c = 0; c < C; ++c) {
    int key = -1, cnt = 0;
    for (int q = q0[c]; q < q1[c]; ++q) {
      const int kq = a.step[q];
      if (kq != key) { key = kq; cnt = 0; }
      for (int b = col_ptr[q] + 1; b < col_ptr[q + 1]; ++b) xcopy_idx[b] = cnt++;
      xcopy_blocks[c] = std::max(xcopy_blocks[c], cnt);
    }
  }
  auto find_block = [&](int row, int col) -> int {
    const int32_t *b0 = blk_row.data() + col_ptr[col], *b1 = blk_row.data() + col_ptr[col + 1];
    const int32_t *it = std::lower_bound(b0, b1, row);
    return (it != b1 && *it == row) ? (int)(it - blk_row.data()) : -1;
  };
  // ---- contribution slots of every CTA != 0: top blocks / top right-hand sides its columns update
  std::vector<std::vector<int>> cb_blk(C), cv_col(C);          // slot -> factor block / column
  std::vector<std::vector<int32_t>> cb_slot_of(C), cv_slot_of(C);
  for (int c = 1; c < C; ++c) { cb_slot_of[c].assign(col_ptr[n] - col_ptr[q0[0]], -1); cv_slot_of[c].assign(n - q0[0], -1); }
  // symbolic destinations, resolved to pool offsets once the slot counts are known
  enum { kOwnBlk = 0, kOwnVec, kCtbBlk, kCtbVec };
  struct Prod { int owner, step, critical, dkind, dkey, a_off, b_off, k, nrows; };  // b_off: index of the unscaled copy
  std::vector<Prod> prods;
  prods.reserve(8 * (size_t)col_ptr[n]);
  const int tb0 = col_ptr[q0[0]];
  for (int c = 0; c < n; ++c) {
    for (int rr = row_ptr[c]; rr < row_ptr[c + 1]; ++rr) {
      const int bck = row_blk[rr], k = row_col[rr];
      const int o = a.cta[k];
      if (!own(o, k)) return fail("internal: source column not owned");
      const int b_off = xcopy_idx[bck];
      const bool dest_own = own(o, c);
      if (!dest_own && (o == 0 || !a.top[c])) return fail("internal: destination neither own nor top");
      {  // right-hand side: r_c -= X_ck y_k
        Prod p{o, a.step[k] + 1, 0, dest_own ? kOwnVec : kCtbVec, c, V0[o] + 6 * (k - q0[o]), b_off, k, 1};
        if (!dest_own) {
          int32_t &sl = cv_slot_of[o][c - q0[0]];
          if (sl < 0) { sl = (int32_t)cv_col[o].size(); cv_col[o].push_back(c); }
          p.dkey = sl;
        }
        prods.push_back(p);
      }
      for (int bik = bck; bik < col_ptr[k + 1]; ++bik) {
        const int i = blk_row[bik];
        const int dst = find_block(i, c);
        if (dst < 0) return fail("internal: symbolic factorisation inconsistent");
        Prod p{o, a.step[k] + 1, 0, dest_own ? kOwnBlk : kCtbBlk, dst, 36 * (bik - tp.b0[o]), b_off, k, 6};
        if (dest_own) {
          if (i == c && a.step[c] == a.step[k] + 1) { p.critical = 1; p.step = a.step[c]; }
        } else {
          int32_t &sl = cb_slot_of[o][dst - tb0];
          if (sl < 0) { sl = (int32_t)cb_blk[o].size(); cb_blk[o].push_back(dst); }
          p.dkey = sl;
        }
        prods.push_back(p);
      }
    }
  }
  std::vector<int> CV0(C), XC0(C);
  for (int c = 0; c < C; ++c) {
    CV0[c] = CB0[c] + 36 * (int)cb_blk[c].size();
    tp.contrib_off[c] = CB0[c];
    tp.contrib_doubles[c] = 36 * (int)cb_blk[c].size() + 6 * (int)cv_col[c].size();
    XC0[c] = CB0[c] + tp.contrib_doubles[c];
    tp.xcopy_off[c] = XC0[c];
    tp.pool_doubles[c] = XC0[c] + 36 * xcopy_blocks[c];
    if (tp.pool_doubles[c] >= 65536) return fail("pool exceeds 16-bit offsets");
  }
  {
    int off = 0;
    for (int c = 0; c < C; ++c) { tp.xchg_off[c] = off; if (c) off += tp.contrib_doubles[c]; }
    tp.xchg_doubles = off;
    if (off >= 65536) return fail("exchange buffer exceeds 15-bit offsets");
  }
  auto dest_off = [&](const Prod &p) {
    switch (p.dkind) {
      case kOwnBlk: return 36 * (p.dkey - tp.b0[p.owner]);
      case kOwnVec: return V0[p.owner] + 6 * (p.dkey - q0[p.owner]);
      case kCtbBlk: return CB0[p.owner] + 36 * p.dkey;
      default: return CV0[p.owner] + 6 * p.dkey;
    }
  };
  {
    // order: owner, step, critical first, destination kind, destination, source column (= the summation order of a
    // destination) - packed into one 64-bit key per product, sorted with its index, then applied
    std::vector<std::pair<uint64_t, uint32_t>> keys(prods.size());
    bool packable = true;
    for (size_t i = 0; i < prods.size(); ++i) {
      const Prod &x = prods[i];
      if (x.owner >= 32 || x.step >= 4096 || x.dkey >= (1 << 24) || x.k >= (1 << 20) || x.step < 0 || x.dkey < 0) { packable = false; break; }
      keys[i] = {((uint64_t)x.owner << 59) | ((uint64_t)x.step << 47) | ((uint64_t)(x.critical ? 0 : 1) << 46) | ((uint64_t)x.dkind << 44) |
                 ((uint64_t)x.dkey << 20) | (uint64_t)x.k, (uint32_t)i};
    }
    if (packable) {
      // bucket by (owner, step) - the top 17 bits of the key - with a counting sort, then sort the (small) buckets
      int max_step = 0;
      for (const Prod &x : prods) max_step = std::max(max_step, x.step);
      const int n_steps = max_step + 1;
      std::vector<uint32_t> bucket_ptr((size_t)C * n_steps + 1, 0);
      for (const Prod &x : prods) ++bucket_ptr[(size_t)x.owner * n_steps + x.step + 1];
      for (size_t b = 0; b + 1 < bucket_ptr.size(); ++b) bucket_ptr[b + 1] += bucket_ptr[b];
      std::vector<std::pair<uint64_t, uint32_t>> scattered(keys.size());
      {
        std::vector<uint32_t> fill(bucket_ptr.begin(), bucket_ptr.end() - 1);
        for (size_t i = 0; i < prods.size(); ++i) scattered[fill[(size_t)prods[i].owner * n_steps + prods[i].step]++] = keys[i];
      }
      for (size_t b = 0; b + 1 < bucket_ptr.size(); ++b)
        if (bucket_ptr[b + 1] - bucket_ptr[b] > 1) std::sort(scattered.begin() + bucket_ptr[b], scattered.begin() + bucket_ptr[b + 1]);
      std::vector<Prod> sorted(prods.size());
      for (size_t i = 0; i < scattered.size(); ++i) sorted[i] = prods[scattered[i].second];
      prods.swap(sorted);
    } else {
      std::sort(prods.begin(), prods.end(), [](const Prod &x, const Prod &y) {
        if (x.owner != y.owner) return x.owner < y.owner;
        if (x.step != y.step) return x.step < y.step;
        if (x.critical != y.critical) return x.critical > y.critical;
        if (x.dkind != y.dkind) return x.dkind < y.dkind;
        if (x.dkey != y.dkey) return x.dkey < y.dkey;
        return x.k < y.k;  // summation order of a destination: by source column
      });
    }
  }
  // ---- emit the program of every CTA
  size_t pi = 0;
  tp.words.clear();
  tp.smem_bytes = 0;
  int chain_a = 0;
  for (int c = 0; c < C; ++c) {
    const int nsa = a.steps_a[c], nsb = c == 0 ? a.steps_b : 0, ns = nsa + nsb;
    if (c) chain_a = std::max(chain_a, nsa); else chain_a = std::max(chain_a, nsa);
    std::vector<int32_t> w(kTH_Words + kTS_Words * (size_t)ns, 0);
    w[kTH_StepsA] = nsa; w[kTH_StepsB] = nsb; w[kTH_OffSteps] = kTH_Words;
    w[kTH_TopCol0] = c == 0 ? tp.n_own_cols[0] - a.n_top : tp.n_own_cols[c];  // first top column (local index)
    // columns of every step, in order
    std::vector<std::vector<int>> step_cols(ns);
    for (int q = q0[c]; q < q1[c]; ++q) {
      if (a.step[q] < 0 || a.step[q] >= ns) return fail("internal: step out of range");
      step_cols[a.step[q]].push_back(q);
    }
    std::vector<std::vector<Item>> pre_of(q1[c] - q0[c]);  // column -> critical panel items feeding it (filled one step earlier)
    for (int s = 0; s < ns; ++s) {
      const std::vector<int> &cols = step_cols[s];
      const int nc = (int)cols.size();
      if (nc > kTreeStepCols) return fail("internal: step too wide");
      // products of this CTA scheduled at this step: critical (per diagonal block) and look-ahead
      std::vector<Item> diag(nc, Item{0, 6, 0, 0}), look;
      std::vector<int32_t> pairs;
      for (int t = 0; t < nc; ++t) diag[t].dest = 36 * (col_ptr[cols[t]] - tp.b0[c]);
      while (pi < prods.size() && prods[pi].owner == c && prods[pi].step == s) {
        const Prod &p0 = prods[pi];
        const int d = dest_off(p0);
        const int pbeg = (int)pairs.size();
        size_t pj = pi;
        while (pj < prods.size() && prods[pj].owner == c && prods[pj].step == s && prods[pj].critical == p0.critical &&
               prods[pj].dkind == p0.dkind && prods[pj].dkey == p0.dkey) {
          pairs.push_back((int32_t)((unsigned)prods[pj].a_off | ((unsigned)(XC0[c] + 36 * prods[pj].b_off) << 16)));
          ++pj;
        }
        if (p0.critical) {
          int t = 0;
          while (t < nc && diag[t].dest != d) ++t;
          if (t == nc) return fail("internal: critical product without its column");
          diag[t].p0 = pbeg; diag[t].p1 = (int)pairs.size();
        } else {
          look.push_back(Item{d, p0.nrows, pbeg, (int)pairs.size()});
        }
        pi = pj;
      }
      // look-ahead rounds: five items of similar length per warp round, longest first; among items of one length the
      // ones that multiply by the same unscaled block X_jk (same source column, same j) side by side, so that the five
      // lane groups of a round load the same B operand (one shared-memory wavefront per load instead of two)
      std::stable_sort(look.begin(), look.end(), [&](const Item &x, const Item &y) {
        const int lx = x.p1 - x.p0, ly = y.p1 - y.p0;
        if (lx != ly) return lx > ly;
        if (lx == 0) return false;
        return ((unsigned)pairs[x.p0] >> 16) < ((unsigned)pairs[y.p0] >> 16);
      });
      int32_t *st = nullptr;
      auto step_entry = [&]() { return w.data() + kTH_Words + kTS_Words * (size_t)s; };
      const int off_pairs = (int)w.size();
      w.insert(w.end(), pairs.begin(), pairs.end());
      if (w.size() & 1) w.push_back(0);  // items are read as 8-byte words
      st = step_entry();
      st[kTS_Cols] = nc;
      st[kTS_OffDiag] = (int)w.size();
      for (int t = 0; t < nc; ++t)
        if (!push_product_item(w, diag[t].dest, 6, off_pairs + diag[t].p0, diag[t].p1 - diag[t].p0)) return fail("too many products for one block");
      st = step_entry();
      st[kTS_NLook] = ((int)look.size() + 4) / 5;
      st[kTS_OffLook] = (int)w.size();
      for (size_t i = 0; i < 5 * (size_t)(((int)look.size() + 4) / 5); ++i) {
        if (i >= look.size()) push_product_item(w, 0, 0, 0, 0);
        else if (!push_product_item(w, look[i].dest, look[i].nrows, off_pairs + look[i].p0, look[i].p1 - look[i].p0)) return fail("too many products for one block");
      }
      // the critical panel items of the previous step, per diagonal item of this one
      st = step_entry();
      st[kTS_OffPre] = (int)w.size();
      {
        const size_t tab = w.size();
        w.resize(w.size() + 2 * (size_t)nc, 0);
        for (int t = 0; t < nc; ++t) {
          std::vector<Item> &pre = pre_of[cols[t] - q0[c]];
          w[tab + 2 * t] = (int)pre.size();
          w[tab + 2 * t + 1] = (int)w.size();
          for (const Item &it : pre) push_panel_item(w, it.dest, it.nrows, it.p0, it.p1);
          pre.clear();
        }
      }
      // panel items: every sub-diagonal block row-wise times M_j, then the right-hand sides.  A block whose ROW is a
      // column of this CTA's next step feeds that column's critical products: it goes to that column's lane group
      std::vector<Item> panel;
      for (int t = 0; t < nc; ++t) {
        const int j = cols[t], dblk = 36 * (col_ptr[j] - tp.b0[c]);
        for (int b = col_ptr[j] + 1; b < col_ptr[j + 1]; ++b) {
          const int i = blk_row[b];
          const Item it{36 * (b - tp.b0[c]), 6, dblk, XC0[c] + 36 * xcopy_idx[b]};
          if (own(c, i) && a.step[i] == s + 1 && s + 1 < ns) pre_of[i - q0[c]].push_back(it);
          else panel.push_back(it);
        }
      }
      for (int t = 0; t < nc; ++t) {
        const int j = cols[t];
        panel.push_back(Item{V0[c] + 6 * (j - q0[c]), 1, 36 * (col_ptr[j] - tp.b0[c]), 0});
      }
      st = step_entry();
      st[kTS_NPanel] = ((int)panel.size() + 4) / 5;
      st[kTS_OffPanel] = (int)w.size();
      for (size_t i = 0; i < 5 * (size_t)(((int)panel.size() + 4) / 5); ++i) {
        if (i < panel.size()) push_panel_item(w, panel[i].dest, panel[i].nrows, panel[i].p0, panel[i].p1);
        else push_panel_item(w, 0, 0, 0, 0);
      }
      // backward rounds of this step: its columns i are the sources, w_j -= Y_ij^T x_i for every own column j
      // with a block (i, j); one item per destination, the destinations of the previous step (the next to become
      // final on the way back) first
      {
        std::vector<BwdPair> bp;
        for (int t = 0; t < nc; ++t) {
          const int i = cols[t];
          for (int rr = row_ptr[i]; rr < row_ptr[i + 1]; ++rr) {
            const int k = row_col[rr];
            if (!own(c, k)) continue;  // a top row reaching into another CTA's subtree: that CTA's top rounds
            bp.push_back(BwdPair{k, a.step[k] == s - 1 ? 0 : 1, i, 36 * (row_blk[rr] - tp.b0[c]), V0[c] + 6 * (i - q0[c])});
          }
        }
        int off_items = 0;
        const int n_rounds = emit_bwd_rounds(w, bp, V0[c], q0[c], off_items);
        if (n_rounds < 0 || n_rounds >= 32768) return fail("too many backward rounds");
        st = step_entry();
        st[kTS_OffBwd] = off_items;
        st[kTS_Cols] = nc | (n_rounds << 16);
      }
    }
    if (c) {
      // the top solution arrives in this CTA's vector slots: its rounds come before the CTA's own backward steps
      std::vector<BwdPair> bp;
      for (int j = q0[c]; j < q1[c]; ++j)
        for (int b = col_ptr[j] + 1; b < col_ptr[j + 1]; ++b) {
          const int i = blk_row[b];
          if (own(c, i)) continue;
          const int sl = a.top[i] ? cv_slot_of[c][i - q0[0]] : -1;
          if (sl < 0) return fail("internal: backward row without a slot");
          bp.push_back(BwdPair{j, -a.step[j], i, 36 * (b - tp.b0[c]), CV0[c] + 6 * sl});
        }
      int off_items = 0;
      const int n_rounds = emit_bwd_rounds(w, bp, V0[c], q0[c], off_items);
      if (n_rounds < 0) return fail("too many backward rounds");
      w[kTH_OffTopBwd] = off_items;
      w[kTH_NTopBwd] = n_rounds;
    }
    if (pi < prods.size() && prods[pi].owner == c) return fail("internal: product scheduled after the last step");
    if (c == 0 && C > 1) {
      // add rounds: round r applies, to every destination, its r-th contribution (sources in CTA order)
      struct Op { int dest, src, is_vec, cta; };
      std::vector<Op> ops;
      for (int cc = 1; cc < C; ++cc) {
        for (size_t sl = 0; sl < cb_blk[cc].size(); ++sl)
          ops.push_back(Op{36 * (cb_blk[cc][sl] - tp.b0[0]), tp.xchg_off[cc] + 36 * (int)sl, 0, cc});
        for (size_t sl = 0; sl < cv_col[cc].size(); ++sl)
          ops.push_back(Op{V0[0] + 6 * (cv_col[cc][sl] - q0[0]), tp.xchg_off[cc] + 36 * (int)cb_blk[cc].size() + 6 * (int)sl, 1, cc});
      }
      std::stable_sort(ops.begin(), ops.end(), [](const Op &x, const Op &y) { return x.dest != y.dest ? x.dest < y.dest : x.cta < y.cta; });
      std::vector<std::vector<int32_t>> rounds;
      for (size_t i = 0; i < ops.size();) {
        size_t j = i;
        while (j < ops.size() && ops[j].dest == ops[i].dest) {
          const size_t r = j - i;
          if (rounds.size() <= r) rounds.emplace_back();
          rounds[r].push_back((int32_t)((unsigned)ops[j].dest | ((unsigned)(ops[j].src / 2) << 16) | (ops[j].is_vec ? 0x80000000u : 0u)));
          ++j;
        }
        i = j;
      }
      w[kTH_AddRounds] = (int)rounds.size();
      w[kTH_OffAddRounds] = (int)w.size();
      const size_t tab = w.size();
      w.resize(w.size() + 2 * rounds.size(), 0);
      for (size_t r = 0; r < rounds.size(); ++r) {
        w[tab + 2 * r] = (int)rounds[r].size();
        w[tab + 2 * r + 1] = (int)w.size();
        w.insert(w.end(), rounds[r].begin(), rounds[r].end());
      }
    } else if (c) {
      // after the top part is solved: x of the top columns this CTA's blocks reach -> their vector slots
      w[kTH_NXload] = (int)cv_col[c].size();
      w[kTH_OffXload] = (int)w.size();
      for (size_t sl = 0; sl < cv_col[c].size(); ++sl) {
        if (cv_col[c][sl] >= 65536) return fail("more than 65535 columns");
        w.push_back((int32_t)((unsigned)(CV0[c] + 6 * (int)sl) | ((unsigned)cv_col[c][sl] << 16)));
      }
    }
    while (w.size() % 4) w.push_back(0);
    tp.prog_ptr[c] = (int32_t)tp.words.size();
    tp.words.insert(tp.words.end(), w.begin(), w.end());
    const size_t bytes = 8 * (size_t)tp.pool_doubles[c] + 4 * w.size() + kTreeMiscBytes + 64;
    tp.smem_bytes = std::max(tp.smem_bytes, bytes);
  }
  for (int c = C; c <= kTreeMaxCluster; ++c) tp.prog_ptr[c] = (int32_t)tp.words.size();
  if (pi != prods.size()) return fail("internal: unscheduled products");
  if (tp.smem_bytes > kTreeMaxSmem) return fail("does not fit the shared memory of the cluster");
  tp.chain_steps = chain_a + a.steps_b;
  tp.ok = true;
  return true;
}

}  // namespace ssba
