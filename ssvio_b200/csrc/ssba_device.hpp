// ssba_device.hpp — device-resident problem (HBM layout) and the kernel launch interface.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "ssba.h"
#include "ssba_geometry.cuh"
#include "ssba_tree_program.hpp"

namespace ssba {

// LM controller state, lives in device memory; the host only reads it back at sync points.
// Mirrors the members of g2o::OptimizationAlgorithmLevenberg
// (g2o/core/optimization_algorithm_levenberg.h:72-81) plus the loop variables of solve()
// (levenberg.cpp:58-150) and of SparseOptimizer::optimize (sparse_optimizer.cpp:386-426).
struct Control {
  // options (levenberg.cpp:44-56)
  double tau, good_lower, good_upper, user_lambda;
  int max_trials;
  // LM state
  double lambda, ni;
  double current_chi, temp_chi, rho;
  double scale_pose_part[kTreeMaxCluster];  // sum over poses of x (lambda x + b): one partial per CTA of the reduced
                           // solve (k_reduced_solve writes [0] only), folded in order by control_step
  double maxdiag;          // max |H_jj| for computeLambdaInit
  double chi2_initial;
  int cur;                 // which of the two state buffers holds the current estimate
  int need_linearize;      // next slot starts a new outer iteration (buildSystem) and k_linearize has to run
  int need_fold;           // the linearisation is new (made by k_update of the accepted trial): fold its Hpp partials
  int lin;                 // which of the two sets of linearisation buffers (W, Hll, b_l, Hpp partials) is current
  int lin_valid;           // buffers `lin` hold the linearisation of the current estimate
  int first_iteration;     // lambda must be initialised (iteration == 0)
  int done;                // optimize() finished: every kernel returns immediately
  int outer_iter, max_iters;
  int qmax;                // _levenbergIterations of the running outer iteration
  int chol_fail;           // the reduced Cholesky of this trial hit a pivot <= 0
  int cholesky_failures;
  int last_result;         // ssba_solver_result of the last finished outer iteration
  int n_records;
  int world, rank;
  unsigned int ticket;     // CTAs of k_update that have finished (the last one runs control_step)
  long long trial_seq;     // trials executed on this handle so far (same on every rank): the sequence number of
                           // the peer-memory exchanges
  int comm_timeout;        // a peer did not publish its part in time (peer-memory exchange)
  int force_stop;          // SparseOptimizer::terminate(): raised through ssba_request_stop (a copy on a side stream)
  long long n_trials;      // levenbergIterations summed over this run
  long long dbg[8];        // SSBA_TIMING: globaltimer stamps of the last trial's exchange kernels
  ssba_iter_record records[SSBA_MAX_ITER_RECORDS];
};

constexpr int SSBA_MAX_PEERS = 16;
// head of a rank's exchange buffer.  Flags and the small partial sums are PUSHED: rank s writes them into
// the `[s]` entries of every peer's header (remote stores over NVLink), so that a rank only ever polls its
// own memory; the partial reduced systems (after the header, 2 x sys_doubles, indexed by trial parity) are
// pulled by the readers.  Flags are trial sequence numbers.
struct PeerHeader {
  long long flag_sys_from[SSBA_MAX_PEERS];    // rank s's partial reduced system of that trial is complete
  long long flag_scal_from[SSBA_MAX_PEERS];   // rank s's partial sums of that trial have arrived
  double scal_from[SSBA_MAX_PEERS][2][4];     // [s][parity]: chi(current), chi(trial), landmark part of computeScale
};

// the subtree-per-CTA solver program on the device (ssba_tree_program.hpp); C == 0: not in use
struct TreeDev {
  int C;
  unsigned smem_bytes;
  const int32_t *prog;
  double *xchg;  // contributions of CTAs 1 .. C-1 to the top part
  int32_t prog_ptr[kTreeMaxCluster + 1];
  int32_t pool_doubles[kTreeMaxCluster], b0[kTreeMaxCluster], n_own_blocks[kTreeMaxCluster], q0[kTreeMaxCluster],
      n_own_cols[kTreeMaxCluster], contrib_off[kTreeMaxCluster], contrib_doubles[kTreeMaxCluster], xchg_off[kTreeMaxCluster];
};

// Everything the per-pair kernels need to know about a (pose, landmark) pair in ONE 32-byte record (two 16-byte
// loads, coalesced over the threads of a chunk) instead of a chain through pair_vertex / pair_q / pair_slot ->
// slot_vertex, slot_free / pair_edge_ptr.  Packed on the device from those arrays when a structure is uploaded.
struct __align__(16) PairRec {
  int32_t pose_row, q, slot, e0;          // q < 0: the pose is fixed; e0: first edge of the pair
  int32_t n_edges, point_row, lfree, pad;  // lfree: the landmark is free
};

constexpr int kReadoutThreads = 128;  // k_final_chi2: one thread per landmark (grid planned by ssba_initialize)

struct DeviceProblem {
  Cameras cams;
  double ext_R[8][9];  // rotation matrices of the extrinsics
  int jacobian_mode;
  double delta_all;
  // sizes
  int n_poses, n_points, n_fp, n_slots, n_pairs, n_edges, n_blocks, n_hpp_parts, n_levels;
  int n_lin_blocks, n_upd_blocks;  // grids of k_linearize / k_update = lengths of their partial-sum arrays
  int n_fin_blocks;                // grid of the thread-per-landmark read-out kernel (kReadoutThreads landmarks per CTA)
  // estimates: two buffers each (current / trial), selected by Control::cur
  double *pose[2];    // n_poses x 7
  double *point[2];   // n_points x 3
  double *pose0, *point0;  // initial estimates (for ssba_reset_state)
  // landmark-major shard
  const int32_t *slot_vertex, *slot_pair_ptr;
  const uint8_t *slot_free;
  const int32_t *pair_vertex, *pair_q, *pair_edge_ptr, *pair_slot, *lchunk_slot;
  const PairRec *pair_rec;  // n_pairs, see PairRec
  const double *e_uv, *e_info, *e_delta;   // e_info / e_delta may be null
  const uint8_t *e_cam;
  const int32_t *e_orig;
  // Hpp partials of the linearize CTAs
  const int32_t *lchunk_lp_ptr, *lp_pair_ptr, *q_part_ptr, *q_part;
  const uint8_t *lp_pair;
  const int32_t *pose_of_q;
  const int32_t *unit_slot, *unit_n, *unit_k, *unit_c0;  // Schur work units
  const int32_t *unit_combo_ptr, *combo_blk;             // ... and the factor blocks they accumulate into
  const int32_t *blk_prod_ptr, *combo_pos;               // per factor block its producers (combos, in unit order) are the
                                                         // staging slots [blk_prod_ptr[b], blk_prod_ptr[b + 1]); combo -> slot
  double *stage, *stage_b;  // deterministic accumulation: per combo its 6x6 total (and, diagonal combos, its 6-vector)
  int deterministic;        // k_schur stores per-unit totals, k_schur_reduce adds them in a fixed order (no fp64 atomics)
  int n_units;
  // factor structure
  const int32_t *blk_row, *blk_col;
  const int32_t *col_diag;         // n_fp: the diagonal block of every column
  const int32_t *prog, *prog_ptr;  // per-level solver program (ssba_structure.cpp build_solver_program)
  int prog_max_seg, n_segments;
  int solve_cluster;               // CTAs the solver program was dealt over (1, 2, 4, 8)
  TreeDev tree;                    // k_tree_solve's program; tree.C == 0: k_reduced_solve and the level program
  // system
  // the linearisation, two sets: Control::lin names the one of the current estimate, k_update writes the trial
  // state's into the other one
  double *W[2];          // n_pairs x 18, 6x3 row-major  (Hpl blocks)
  double *Hll[2];        // n_slots x 6 (xx xy xz yy yz zz)
  double *bl[2];         // n_slots x 3
  double *hpp_part[2];   // n_hpp_parts x 27 (b[6], upper-tri H[21])
  double *Dinv;       // n_slots x 6
  double *hpp_fold;   // n_fp x 27: the partials of each pose folded (k_prepare_system)
  double *sys;        // [ L: n_blocks x 36 | bschur: n_fp x 6 | bp: n_fp x 6 ] — the all-reduced buffer
  double *xp;         // n_fp x 6
  double *diag_buf;   // n_fp x 6 Hpp diagonals (lambda init, summed over ranks)
  double *chi_cur_part, *maxdiag_part;      // n_lin_blocks
  double *chi_new_part, *scale_part;        // n_upd_blocks
  double *scal;       // [chi_cur, chi_new, scale_lm, maxdiag] reduced partials (all-reduced)
  double *err_out;    // n_edges_total x 2 scratch for ssba_get_edge_errors
  uint8_t *mask_out;  // n_edges_total: ssba_get_outlier_mask
  double *chi_out;    // [plain, robust, n_outliers, n_inliers]
  double *gather;     // n_points x 3, multi-GPU read-back of the landmark estimates
  const uint8_t *owner_mask;  // n_points: this rank reports the landmark
  Control *ctl;
  // several GPUs, peer-memory exchange (ssba_api.cu setup_peer_exchange): every rank's exchange buffer as
  // mapped into this process; layout per rank: PeerHeader, then 2 x sys_doubles partial reduced systems
  char *peer[SSBA_MAX_PEERS];
  int use_p2p;
  unsigned pose_stage;  // bytes of the pose array the edge kernels stage into shared memory with one bulk copy (0: gather from L2)
  int pdl;  // launch k_schur / the reduced solve / k_update with programmatic dependent launch (off while profiling)
  size_t sys_doubles;
  int n_edges_total;
};

// Programmatic dependent launch (griddepcontrol): a kernel launched with the attribute may start while its
// predecessor in the stream still runs; everything before griddep_wait() must only touch STATIC data (the index
// structure), griddep_wait() returns once the predecessor has completed and its writes are visible.
// griddep_launch() lets the successor's CTAs be scheduled from here on.  Both are no-ops in a plain launch.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
template <class... KArgs, class... Args>
inline cudaError_t launch_maybe_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}

// launches (all asynchronous on `st`)
void launch_linearize(const DeviceProblem &P, cudaStream_t st);        // K_lin + K_hpp + hpp reduce
void launch_maxdiag(const DeviceProblem &P, cudaStream_t st);          // -> scal[3]
void launch_lambda_init(const DeviceProblem &P, cudaStream_t st);
// sys += blockdiag(Hpp) + lambda I - W Hll^-1 W^T, bschur, b_p (sys zeroed by the previous k_update)
void launch_schur(const DeviceProblem &P, bool prefolded, cudaStream_t st);  // + k_schur_reduce when P.deterministic
void launch_reduced_solve(const DeviceProblem &P, cudaStream_t st);
void launch_update(const DeviceProblem &P, bool fused_control, cudaStream_t st);  // + accept/reject when fused
bool update_linearizes(const DeviceProblem &P);
unsigned pose_stage_bytes(int n_poses);  // what fits beside the linearisation scratch, else 0  // k_update also linearises the trial state (closed-form Jacobians)
void launch_reduce_partials(const DeviceProblem &P, cudaStream_t st);  // -> scal[0..2]
void launch_control(const DeviceProblem &P, cudaStream_t st);          // several GPUs only
void launch_exchange_sys(const DeviceProblem &P, cudaStream_t st);     // peer-memory all-reduce of the reduced system
void launch_control_p2p(const DeviceProblem &P, cudaStream_t st);      // partial sums exchanged through peer memory + decision
void launch_fold(const DeviceProblem &P, cudaStream_t st);           // first slot: hpp_fold + Hpp diagonals
void launch_final_chi2(const DeviceProblem &P, double threshold, cudaStream_t st);  // -> chi_out
void launch_edge_errors(const DeviceProblem &P, cudaStream_t st);      // -> err_out
void launch_outlier_mask(const DeviceProblem &P, double threshold, cudaStream_t st);  // -> mask_out, chi_out[2] = #outliers
void launch_gather_points(const DeviceProblem &P, cudaStream_t st);    // -> gather
void launch_pack_pairs(const DeviceProblem &P, cudaStream_t st);       // pair_* / slot_* -> pair_rec
// uv / information / Huber widths from the caller's edge order (raw_*) into the structure's order (P.e_*)
void launch_gather_edge_values(const DeviceProblem &P, const double *raw_uv, const double *raw_info, const double *raw_delta, cudaStream_t st);
int kernels_per_linearize();

// pose-graph optimisation: one LM trial (zero / linearise when needed, lambda_0 in the first slot, H + lambda I,
// the level-scheduled solver, chi2 of current and trial poses, accept / reject); chi2 of the current poses
void launch_pose_graph_slot(const DeviceProblem &P, int n_edges, const int32_t *ev0, const int32_t *ev1, const int32_t *eq0,
                            const int32_t *eq1, const int32_t *eblk, const double *minv, double *sysH, double *stage,
                            const int32_t *prod_ptr, const int32_t *prod, bool first, cudaStream_t st);
void launch_pose_graph_chi(const DeviceProblem &P, int n_edges, const int32_t *ev0, const int32_t *ev1, const double *minv,
                           cudaStream_t st);
int max_solver_cluster();  // largest k_reduced_solve cluster the current device can co-schedule
int max_tree_cluster();    // ... and the largest k_tree_solve cluster (16 when the non-portable size is available)
void launch_tree_solve(const DeviceProblem &P, cudaStream_t st);  // ssba_tree_solve.cu
// host: fills TreeDev from a planned program (prog / xchg are device pointers)
void fill_tree_dev(const TreeProgram &tp, const int32_t *d_prog, double *d_xchg, TreeDev &out);

// batched pose-only LM (ssba_pose_only.cu): one warp per frame, everything in one launch
void launch_pose_only(const double K[9], int n_frames, int rounds, int iters, int pre_rounds, int max_trials, double chi2_threshold,
                      double tau, double good_lower, double good_upper, double user_lambda, const int32_t *feat_ptr, const double *poses_in,
                      const double *xyz, const double *uv, double *err, uint8_t *outlier, double *poses_out,
                      double *chi2_out, int32_t *n_inliers_out, cudaStream_t st);

}  // namespace ssba
