// ssba_structure.hpp — host-side structure of one local-BA problem: what g2o builds in
// SparseOptimizer::initializeOptimization (g2o/core/sparse_optimizer.cpp:168-272),
// BlockSolver::buildStructure (g2o/core/block_solver.hpp:102-256) and
// LinearSolverCSparse::computeSymbolicDecomposition (g2o/solvers/csparse/linear_solver_csparse.h:
// 246-308), flattened into the index arrays the CUDA kernels walk.
//
// Vocabulary
//   slot   an active landmark of THIS rank's shard (free or fixed), landmark-major order
//   pair   all edges joining one (pose, landmark) — one Hpl block W (6x3) when both are free
//   q      permuted index of a free pose = its block row/column in the reduced (Schur) system
//   block  a 6x6 block of the lower-triangular factor L (CSC over q, diagonal first)
#pragma once

#include <cstdint>
#include <functional>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "ssba_geometry.cuh"
#include "ssba_tree_program.hpp"

namespace ssba {

// std::vector whose resize() leaves new elements uninitialised: the big per-edge / per-pair arrays
// are filled completely (in parallel) right after they are sized, a zero-fill first would be a
// second pass over several megabytes on the end-to-end path
template <class T>
struct DefaultInitAllocator : std::allocator<T> {
  template <class U> struct rebind { using other = DefaultInitAllocator<U>; };
  using std::allocator<T>::allocator;
  template <class U> void construct(U *p) noexcept { ::new (static_cast<void *>(p)) U; }
  template <class U, class... A> void construct(U *p, A &&...a) { ::new (static_cast<void *>(p)) U(std::forward<A>(a)...); }
};
template <class T> using uvec = std::vector<T, DefaultInitAllocator<T>>;

struct HostGraph {
  Cameras cams{};
  bool have_cams = false;
  int n_poses = 0, n_points = 0, n_edges = 0;
  std::vector<double> poses, points;           // 7 / 3 doubles per vertex
  std::vector<uint8_t> pose_fixed, point_fixed;
  uvec<int32_t> e_pose, e_point;
  uvec<uint8_t> e_cam;
  uvec<double> e_uv;                           // 2 per edge
  uvec<double> e_info;                         // 3 per edge or empty (identity)
  uvec<double> e_delta;                        // 1 per edge or empty (delta_all)
  double delta_all = 0.0;
  // the measurement values (uv, information, Huber widths) never visit these vectors: the caller's arrays went to
  // the device in the caller's order and a kernel gathers them into the structure's edge order (ssba_api.cu);
  // build_structure then fills indices only
  bool values_on_device = false, has_info = false, has_delta = false;
};

struct Structure {
  // ---- global (identical on every rank)
  int n_fp = 0;                       // free active poses
  int n_fl_global = 0;                // free active landmarks, all ranks
  int n_active_edges_global = 0;
  long long n_inactive_edges_global = 0;  // edges whose two ends are fixed (never active), all ranks
  std::vector<int32_t> q_of_pose;     // pose row -> q, -1 fixed / inactive
  std::vector<int32_t> pose_of_q;     // q -> pose row
  std::vector<uint8_t> point_active;  // point row has an active edge (on any rank)
  // ---- this rank's shard, landmark-major
  int n_slots = 0, n_pairs = 0, n_edges = 0, n_fl = 0;
  std::vector<int32_t> slot_vertex;   // point row
  std::vector<uint8_t> slot_free;
  std::vector<int32_t> slot_pair_ptr; // n_slots + 1
  uvec<int32_t> pair_vertex;   // pose row
  uvec<int32_t> pair_q;        // -1: pose fixed (pairs with a free pose first, by q)
  uvec<int32_t> pair_edge_ptr; // n_pairs + 1
  uvec<int32_t> pair_slot;     // n_pairs: the slot of the pair
  std::vector<int32_t> lchunk_slot;   // n_lchunks + 1: CTAs of k_linearize / k_update = runs of whole
                                      // landmarks with <= 128 pairs (a larger landmark is alone)
  int n_lchunks = 0;
  // Hpp partials: every linearize CTA folds the pose-side quadratic forms of its pairs per
  // distinct free pose ("local pose") and writes one 27-vector per (chunk, local pose)
  int n_hpp_parts = 0;
  std::vector<int32_t> lchunk_lp_ptr;  // n_lchunks + 1 -> first partial (= local pose) of the chunk
  std::vector<int32_t> lp_pair_ptr;    // n_hpp_parts + 1 -> lp_pair
  std::vector<uint8_t> lp_pair;        // pair index inside the chunk (chunks with <= 128 pairs)
  std::vector<int32_t> q_part_ptr;     // n_fp + 1 -> q_part: the partials of pose q, in chunk order
  std::vector<int32_t> q_part;
  uvec<double> e_uv;                  // sorted copies
  uvec<uint8_t> e_cam;
  uvec<int32_t> e_orig;               // index in the caller's addEdge order
  uvec<double> e_info, e_delta;
  // Schur work units (k_schur): run of landmarks [unit_slot, +unit_n) sharing one W pose list of
  // unit_k poses, block pairs [unit_c0, +32) of its k(k+1)/2
  int n_units = 0;
  std::vector<int32_t> unit_slot, unit_n, unit_k, unit_c0;
  // Schur accumulation targets of every unit: for its W-pairs a <= b (sorted by q), block (q_b, q_a)
  std::vector<int32_t> unit_combo_ptr;  // n_units + 1
  uvec<int32_t> combo_blk;
  // ... and per factor block the combos that produce it, in unit order: the fixed summation order of the
  // deterministic accumulation (k_schur_reduce)
  std::vector<int32_t> blk_prod_ptr;    // n_blocks + 1
  std::vector<int32_t> blk_prod;        // combo indices
  std::vector<int32_t> combo_pos;       // combo -> its position in blk_prod: where the unit stores its total, so that
                                        // the producers of a block are contiguous in memory
  // ---- reduced system: lower block-CSC factor pattern (with fill) over q
  int n_blocks = 0, n_schur_blocks = 0;
  std::vector<int32_t> col_ptr;       // n_fp + 1, diagonal block first in every column
  std::vector<int32_t> blk_row;       // n_blocks
  std::vector<int32_t> blk_col;       // n_blocks
  // work items of the device solver, grouped by level (diagonal items first):
  //   task_dst >= 0 : L[dst] <- (A[dst] - sum over its pairs of L[pair_a] L[pair_b]^T) L(j,j)^-T
  //   task_dst <  0 : forward substitution of column j = -1 - task_dst
  int n_tasks = 0;
  std::vector<int32_t> ltask_ptr;     // n_levels + 1 -> tasks
  std::vector<int32_t> task_dst;      // n_tasks
  std::vector<int32_t> task_pos;      // n_tasks: position of the item's column inside its level
  std::vector<int32_t> task_pair_ptr; // n_tasks + 1 -> pairs
  std::vector<int32_t> pair_a, pair_b;
  double est_solver_cycles = 0.0;     // cost model that picked the elimination order
  double seconds_symbolic = 0.0;      // host time of ordering + symbolic factorisation + solver program
  // the same schedule packed level by level for the device (see build_solver_program)
  std::vector<int32_t> prog, prog_ptr;
  int prog_max_seg = 0;               // ints in the largest level segment
  int n_segments = 0;                 // n_levels + 1 (prologue)
  int solve_cluster = 1;              // CTAs the rounds of every level are dealt over
  int solver_slots = 0, solver_cached_blocks = 0;  // shared-memory factor cache plan
  int solver_rounds = 0;              // warp rounds of the factorisation
  std::vector<int32_t> row_ptr;       // n_fp + 1: strictly-lower blocks by row (forward solve)
  std::vector<int32_t> row_blk, row_col;
  std::vector<int32_t> level_ptr, level_col; // columns grouped by elimination-tree level
  int n_levels = 0;
  // original-order map for error read-back
  int n_edges_total = 0;
  // the subtree-per-CTA solver program (k_tree_solve); when !tree.ok the level program above is used
  TreeProgram tree;
};

// Builds the structure for `rank` of `world`. Returns false and sets err on invalid input.
// n_fp + n_fl_global == 0 is not an error here (caller maps it to SSBA_ERR_EMPTY).
// `on_edges_ready` (optional) is called once the per-edge / per-pair arrays (slot_*, pair_*, e_*)
// are final, before the solver program and the small index lists are built.
//
// `across_ranks` (optional) switches to PRE-SHARDED input: `g` holds only the edges of the landmarks this rank owns
// (all of them), poses and fixed flags are the same on every rank.  Every landmark with a local edge becomes a
// slot (rank / world are not used to select), and what has to be the same on every rank - which poses are active,
// the co-visibility pattern of the reduced system, the global counts - is agreed through the callback:
// across_ranks(bytes, n_bytes, sums, n_sums) replaces `bytes` by the element-wise maximum and `sums` by the sum
// over the ranks (returns false on a communication failure).  Ordering, symbolic factorisation and the solver
// program then come out identical on every rank because they only depend on that pattern.
using AcrossRanks = std::function<bool(uint8_t *bytes, size_t n_bytes, long long *sums, int n_sums)>;
bool build_structure(const HostGraph &g, int rank, int world, Structure &s, std::string &err,
                     const std::function<void()> *on_edges_ready = nullptr, const AcrossRanks *across_ranks = nullptr);

// Solver-only structure for a block-sparse SPD system over n 6x6 block columns (see ssba_structure.cpp).
bool build_solver_structure(int n, const std::vector<std::vector<int>> &adj, Structure &s, std::vector<int> &perm_out,
                            std::string &err);

// How many ranks (processes) share this host: sizes the host thread pool; call before the first build.
void set_ranks_on_host(int n);

// memcpy of several regions on the host thread pool of the structure builder
struct CopyJob { void *dst; const void *src; size_t bytes; };
void parallel_copy(const std::vector<CopyJob> &jobs);
bool parallel_equal(const std::vector<CopyJob> &jobs);  // every job: dst[0, bytes) == src[0, bytes)
void parallel_gather_doubles(double *dst, const double *src, const int32_t *idx, size_t n, int k);

// Which rank owns which landmark under the sharding rule of build_structure:
// owner[point row] = rank, or -1 for landmarks without an active edge.
bool plan_shards(const HostGraph &g, int world, std::vector<int32_t> &owner, std::string &err);

}  // namespace ssba
