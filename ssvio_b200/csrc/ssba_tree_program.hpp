// ssba_tree_program.hpp — host-planned program of k_tree_solve, the reduced pose solve
// (replaces LinearSolverCSparse::solve, g2o/solvers/csparse/linear_solver_csparse.h:106-142 +
// cs_chol_workspace / cs_lsolve / cs_ltsolve, csparse_extension.cpp:35-122).
//
// Idea: the elimination tree of the (nested-dissection ordered) reduced system is cut into a TOP
// part (the upper separators) and disjoint SUBTREES.  Every CTA of a thread-block cluster owns a set
// of subtrees and factors them completely out of its own shared memory, with no communication at
// all; what its columns contribute to the top part is accumulated locally ("contribution blocks")
// and shipped once.  CTA 0 then adds the contributions in CTA order (deterministic), factors the top
// part, solves it, and sends the top solution back down; every CTA finishes the backward substitution
// of its subtrees on its own.  Two cluster barriers per solve instead of one per elimination level.
//
// Inside a CTA the columns are processed in STEPS (a step = the columns of one elimination level of
// that CTA's forest).  Numerics per column j: block LDL^T with 6x6 blocks (the same factorisation as the
// scalar Cholesky of CSparse, grouped differently; D_j is positive definite iff its six Cholesky pivots are
// positive, csparse_extension.cpp:115):
//   D_j  = A_jj - sum_k Y_jk X_jk^T,   M_j = D_j^-1            (closed form, one lane: no chain of six square roots)
//   X_ij = A_ij - sum_k Y_ik X_jk^T,   Y_ij = X_ij M_j         (one lane per block row; Y replaces X in place, the
//                                                               unscaled X_ij goes to a scratch copy for one step)
//   z_j  = b_j - sum_k X_jk w_k,       w_j = M_j z_j           (the right-hand side is one more block row of height 1)
// and backwards  x_j = w_j - sum_{i>j} Y_ij^T x_i  (no triangular solve on the way back), right-looking: as soon as
// the columns i of a step are final, every w_j they reach is updated (one lane group per destination j), so that the
// chain from one step to the next is a single 6x6 product - the other terms of a column were subtracted earlier.
// The products of a finished column k are applied eagerly (right-looking), ALL of them in the step after k's
// (which is why the unscaled copy only lives for one step): the ones a diagonal block of the next step waits
// for by the warp that inverts it, all others by the remaining warps WHILE the diagonal blocks are being
// inverted.  ONE __syncthreads per step: the scaling Y = X M of a finished column's blocks happens at the start of
// the next step - the blocks the next diagonal blocks wait for by their own lane groups, the rest by the
// look-ahead warps, which meet the diagonal warps' results at a named barrier before their products.
//
// Everything the kernel walks is planned here; a CPU interpreter of the same program
// (tests/cpp/test_structure.cpp) checks it against a dense solve.
#pragma once

#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace ssba {

constexpr int kTreeThreads = 384;
constexpr int kTreeWarps = kTreeThreads / 32;
constexpr int kTreeMaxCluster = 16;  // 8 is the portable cluster size, 16 needs cudaFuncAttributeNonPortableClusterSizeAllowed
constexpr size_t kTreeMaxSmem = 227 * 1024;       // opt-in dynamic shared memory per CTA on sm_100
constexpr size_t kTreeMiscBytes = 192;            // mbarrier, round counter, flags, reduction scratch
constexpr int kTreeStepCols = 5 * kTreeWarps;     // diagonal blocks one step can factor (5 lane groups per warp)

// A work item is two words: { dest | nrows << 16 | n_pairs << 20,  first pair word } for product items,
// { dest | nrows << 16, diagonal block | unscaled copy << 16 } for panel items (offsets in doubles into the CTA's
// pool); a pair word is  a | b << 16  (a: rows of a scaled block Y or a vector w, b: an unscaled copy X).
// Rounds are five items.
constexpr int kTreeItemWords = 2, kTreeRoundWords = 5 * kTreeItemWords;
// program header words
// kTH_NTopBwd / kTH_OffTopBwd (CTAs != 0): backward rounds whose sources are the top columns (their solution arrives
// from CTA 0), run before the CTA's own backward steps
enum : int { kTH_StepsA = 0, kTH_StepsB, kTH_OffSteps, kTH_AddRounds, kTH_OffAddRounds, kTH_NXload, kTH_OffXload, kTH_TopCol0, kTH_NTopBwd, kTH_OffTopBwd, kTH_Words = 16 };
// step table entry (8 words)
// kTS_OffPre: per diagonal item of the step two words { n, offset } = the panel items of the PREVIOUS step's columns
// that feed this column's critical products (its own block row): the lane group that inverts the column scales them
// itself, first thing in the step; kTS_NPanel / kTS_OffPanel: the other panel items of THIS step's columns, run by the
// look-ahead warps at the start of the next step (or after the last one)
// kTS_Cols: columns | backward rounds << 16; kTS_OffBwd: the backward rounds of the step = five items
// { destination vector | n_pairs << 16, first pair word } each, pair word = block Y_ij | x_i << 16: w_j -= Y_ij^T x_i for
// the columns i of THIS step (final when the rounds run) and every own column j they reach; the destinations in the
// step before come first
enum : int { kTS_Cols = 0, kTS_OffDiag, kTS_NLook, kTS_OffLook, kTS_NPanel, kTS_OffPanel, kTS_OffBwd, kTS_OffPre, kTS_Words = 8 };

struct TreeProgram {
  bool ok = false;
  int C = 1;                          // CTAs of the cluster
  std::vector<int32_t> words;         // the programs of all CTAs, each padded to a multiple of 4 words
  int32_t prog_ptr[kTreeMaxCluster + 1] = {0};
  // shared-memory pool of CTA c (doubles): [own factor blocks | own rhs/solution vectors | b_p copy |
  // contribution blocks | contribution vectors | unscaled copies of one step]
  int32_t pool_doubles[kTreeMaxCluster] = {0};
  int32_t b0[kTreeMaxCluster] = {0}, n_own_blocks[kTreeMaxCluster] = {0};  // first factor block / count
  int32_t q0[kTreeMaxCluster] = {0}, n_own_cols[kTreeMaxCluster] = {0};    // first column / count (CTA 0: incl. the top part)
  int32_t contrib_off[kTreeMaxCluster] = {0}, contrib_doubles[kTreeMaxCluster] = {0};
  int32_t xcopy_off[kTreeMaxCluster] = {0};  // scratch: unscaled copies of one step's sub-diagonal blocks
  int32_t xchg_off[kTreeMaxCluster] = {0};  // where CTA c's contributions go in the exchange buffer (doubles)
  int32_t xchg_doubles = 0;
  size_t smem_bytes = 0;              // dynamic shared memory of the launch (max over the CTAs)
  int chain_steps = 0;                // steps on the critical path (longest subtree + top)
  int n_top_cols = 0;
  std::string why_not;                // when !ok
};

// Which CTA factors which column in which step (indexed by the NEW column order; order[new] = old).
struct TreeAssign {
  int C = 1;
  std::vector<int> order, cta, step;
  std::vector<uint8_t> top;
  int steps_a[kTreeMaxCluster] = {0};  // steps of the subtree phase of every CTA, incl. its flush step
  int steps_b = 0;                     // steps of the top part (CTA 0)
  int n_top = 0;
};
// cta0_subtree: CTA 0 also factors subtrees (else only the top part: more shared memory for it)
void tree_assign(int n, const std::vector<int32_t> &col_ptr, const std::vector<int32_t> &blk_row, int C_want, bool cta0_subtree,
                 TreeAssign &a);
// largest cluster k_tree_solve may use on this device (ssba_create asks the device once; 8 until then)
void set_tree_cluster_cap(int cap);
int tree_cluster_cap();
size_t tree_smem_estimate(int n, const std::vector<int32_t> &col_ptr, const std::vector<int32_t> &blk_row, const TreeAssign &a);
// col_ptr .. row_col: the symbolic factor under the order of tree_assign (Structure fields of the same names)
bool build_tree_program(int n, const std::vector<int32_t> &col_ptr, const std::vector<int32_t> &blk_row,
                        const std::vector<int32_t> &row_ptr, const std::vector<int32_t> &row_blk,
                        const std::vector<int32_t> &row_col, const TreeAssign &a, TreeProgram &tp);

}  // namespace ssba
