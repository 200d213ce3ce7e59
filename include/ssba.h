/*
 * ssba.h — C ABI of libssba: B200-native (sm_100a) local bundle adjustment for ssvio.
 *
 * This is the drop-in boundary for ONE hot path of weihaoysgs/ssvio: the sparse
 * Levenberg-Marquardt solve that Backend::OptimizeActiveMap() runs through g2o
 * (reference: src/ssvio/backend.cpp:78-203).  Every entry point below names the
 * reference interface it replaces (file:line, paths relative to the ssvio tree;
 * "g2o/" = thirdparty/g2o/g2o/).
 *
 * Conventions
 *   - plain C, no C++ types, no exceptions, no torch types; all pointers are HOST
 *     pointers owned by the caller unless the name ends in _device;
 *   - every function returns an ssba_status; ssba_last_error() gives the message;
 *   - a pose is 7 doubles  (qx, qy, qz, qw, tx, ty, tz)  = Sophus::SE3d memory order
 *     (unit quaternion in Eigen coefficient order, then translation) — T_cw, the
 *     estimate of ssvio::VertexPose (include/ssvio/g2otypes.hpp:28-46);
 *   - a point is 3 doubles = the estimate of ssvio::VertexXYZ (g2otypes.hpp:49-65);
 *   - an edge is one ssvio::EdgeProjection (g2otypes.hpp:112-162): 2-D pixel
 *     measurement, pose index, point index, camera index (which extrinsic);
 *   - one handle = one g2o::SparseOptimizer + OptimizationAlgorithmLevenberg +
 *     BlockSolver_6_3 + LinearSolverCSparse instance (backend.cpp:81-86).  Handles
 *     are independent and re-entrant across threads (one caller at a time per
 *     handle), as the reference's stack-allocated optimizers are.
 *   - there is NO CPU fallback: without a usable CUDA device ssba_create() fails
 *     with SSBA_ERR_NO_DEVICE.
 */
#ifndef SSBA_H
#define SSBA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSBA_VERSION_MAJOR 0
#define SSBA_VERSION_MINOR 3  /* 0.2: + ssba_pose_only_optimize, ssba_pose_graph_optimize, ssba_set_profiling;
                                0.3: + ssba_drop_structure (structure reuse), ssba_pose_only_optimize_loop, solver fields of
                                ssba_problem_info */

#define SSBA_MAX_ITER_RECORDS 128
#define SSBA_MAX_CAMERAS 8
#define SSBA_NCCL_ID_BYTES 128

typedef struct ssba_handle ssba_handle;

typedef enum ssba_status {
  SSBA_OK = 0,
  SSBA_ERR_INVALID_ARG = 1,
  SSBA_ERR_CUDA = 2,
  SSBA_ERR_NO_DEVICE = 3,
  SSBA_ERR_STATE = 4,      /* call order violated (e.g. optimize before set_edges) */
  SSBA_ERR_NCCL = 5,
  SSBA_ERR_EMPTY = 6,      /* nothing to optimise: SparseOptimizer::optimize() returns -1
                              (g2o/core/sparse_optimizer.cpp:368-371) */
  SSBA_ERR_ALLOC = 7
} ssba_status;

/* g2o::OptimizationAlgorithm::SolverResult (g2o/core/optimization_algorithm.h:55) */
typedef enum ssba_solver_result {
  SSBA_SOLVER_TERMINATE = 2,
  SSBA_SOLVER_OK = 1,
  SSBA_SOLVER_FAIL = -1
} ssba_solver_result;

typedef enum ssba_jacobian_mode {
  SSBA_JACOBIAN_ANALYTIC = 0, /* closed form (corrected for the right camera), default   */
  SSBA_JACOBIAN_NUMERIC = 1   /* central differences, delta = 1e-9, the reference as shipped
                                 (g2o/core/base_binary_edge.hpp:144-212)                    */
} ssba_jacobian_mode;

/*
 * Options.  The Levenberg constants mirror the constructor of
 * g2o::OptimizationAlgorithmLevenberg (g2o/core/optimization_algorithm_levenberg.cpp:44-56)
 * and its two g2o::Property overrides "initialLambda" / "maxTrialsAfterFailure".
 */
typedef struct ssba_options {
  double tau;                       /* 1e-5  (:48)                                   */
  double good_step_lower_scale;     /* 1/3   (:50)                                   */
  double good_step_upper_scale;     /* 2/3   (:49)                                   */
  double user_lambda_init;          /* 0 => tau * max|H_jj| (:152-166)               */
  int32_t max_trials_after_failure; /* 10    (:52)                                   */
  int32_t jacobian_mode;            /* ssba_jacobian_mode                            */
  int32_t device_id;                /* CUDA device ordinal, -1 => current device     */
  int32_t profile;                  /* 1 => per-phase CUDA-event timings are kept    */
  void *stream;                     /* cudaStream_t to launch on; NULL => own stream */
  /* multi-GPU: landmarks (with their edges) are sharded over `world_size` handles,
   * one per process/GPU; the reduced pose system is all-reduced over NCCL. */
  int32_t rank;                     /* 0..world_size-1                               */
  int32_t world_size;               /* 1 => single GPU                               */
  uint8_t nccl_id[SSBA_NCCL_ID_BYTES]; /* from ssba_nccl_unique_id() on rank 0       */
  int32_t presharded;               /* world_size > 1 only.  0: every rank is handed the WHOLE graph and keeps the
                                     * landmarks ssba_plan_shards assigns to it.  1: every rank is handed only the
                                     * edges of the landmarks it owns (any partition by landmark; all edges of a
                                     * landmark on one rank), the same poses / fixed flags / n_points everywhere; point
                                     * rows of other ranks' landmarks are not read.  The ranks agree on the active
                                     * poses and on the co-visibility pattern through NCCL inside ssba_initialize
                                     * (collective: all ranks must call it); per-edge read-outs
                                     * (ssba_get_edge_errors, ssba_get_outlier_mask) then cover the rank's own edges */
  int32_t reserved[7];
} ssba_options;

/* One outer LM iteration = one OptimizationAlgorithmLevenberg::solve() call
 * (levenberg.cpp:58-150); what `optimizer.setVerbose(true)` prints per iteration
 * (g2o/core/sparse_optimizer.cpp:411-423, levenberg.cpp:187-193). */
typedef struct ssba_iter_record {
  double chi2;     /* activeRobustChi2() after the iteration                          */
  double lambda;   /* currentLambda() after the iteration                             */
  int32_t trials;  /* levenbergIter: inner trials used (1 = first step accepted)      */
  int32_t result;  /* ssba_solver_result of this iteration                            */
} ssba_iter_record;

typedef struct ssba_report {
  int32_t iterations;      /* return value of SparseOptimizer::optimize()
                              (sparse_optimizer.cpp:427-430): iterations run, 0 on Fail,
                              -1 when there was nothing to optimise                     */
  int32_t last_result;     /* ssba_solver_result of the last iteration                  */
  int32_t n_records;       /* valid entries in iters[]                                  */
  int32_t cholesky_failures; /* trials rejected because a pivot was <= 0
                              (g2o/solvers/csparse/csparse_extension.cpp:115)           */
  double chi2_initial;     /* activeRobustChi2() before the first iteration             */
  double chi2_robust;      /* activeRobustChi2() at the final estimate                  */
  double chi2_plain;       /* activeChi2() at the final estimate                        */
  double lambda;           /* final currentLambda()                                     */
  double seconds_total;    /* host wall time of this call                               */
  double seconds_setup;    /* of which: structure build + upload (buildStructure part)  */
  ssba_iter_record iters[SSBA_MAX_ITER_RECORDS];
} ssba_report;

/* Per-phase device time (CUDA events on the launch stream), accumulated since the
 * last ssba_profile_reset(); the names follow g2o::G2OBatchStatistics
 * (g2o/core/batch_stats.h:40-77). Only filled when options.profile != 0. */
typedef struct ssba_profile {
  double ms_linearize;      /* timeResiduals + timeQuadraticForm                       */
  double ms_schur;          /* timeSchurComplement                                     */
  double ms_reduced_solve;  /* timeLinearSolver                                        */
  double ms_update_chi2;    /* back-substitution + timeUpdate + trial residuals        */
  double ms_allreduce;      /* multi-GPU only                                          */
  int64_t n_linearize, n_schur, n_reduced_solve, n_update_chi2, n_allreduce;
  int64_t kernel_launches;  /* kernels launched by libssba since the reset             */
  /* the remaining G2OBatchStatistics fields (always filled, profile on or off) */
  int64_t levenberg_iterations;      /* levenbergIterations summed over the outer iterations since the reset */
  int64_t outer_iterations;          /* solve() calls (outer iterations) since the reset                      */
  int64_t cholesky_nnz;              /* choleskyNNZ: scalar non-zeros of the factor of the reduced system     */
  int32_t hessian_pose_dimension;    /* hessianPoseDimension = 6 x free poses                                 */
  int32_t hessian_landmark_dimension;/* hessianLandmarkDimension = 3 x free landmarks                         */
  double ms_symbolic_decomposition;  /* timeSymbolicDecomposition: host, ordering + symbolic factorisation +  */
                                     /* solver program of the last structure build                             */
  double ms_numeric_decomposition;   /* timeNumericDecomposition + timeLinearSolution = ms_reduced_solve: the  */
                                     /* device solver factors and substitutes in one kernel                    */
  double ms_structure_build;         /* buildStructure equivalent: host wall time of the last structure build  */
} ssba_profile;

/* ---- life cycle ------------------------------------------------------------------ */

void ssba_default_options(ssba_options *opt);

/* Replaces backend.cpp:81-86 (BlockSolver_6_3 + LinearSolverCSparse +
 * OptimizationAlgorithmLevenberg + SparseOptimizer::setAlgorithm). */
ssba_status ssba_create(const ssba_options *opt, ssba_handle **out);

/* ~SparseOptimizer (g2o/core/sparse_optimizer.cpp:57-61). NULL is allowed. */
void ssba_destroy(ssba_handle *h);

/* Message of the last failing call on this handle (or on create when h == NULL). */
const char *ssba_last_error(const ssba_handle *h);

/* rank 0 calls this and ships the bytes to the other ranks (e.g. torch.distributed
 * broadcast); every rank then passes them in ssba_options.nccl_id. */
ssba_status ssba_nccl_unique_id(uint8_t out[SSBA_NCCL_ID_BYTES]);

/* ---- graph construction (backend.cpp:88-168) --------------------------------------- */

/* Camera intrinsics and extrinsics: the (K, cam_ext) pair every EdgeProjection is
 * constructed with (g2otypes.hpp:118-121; backend.cpp:105-107,147-155).
 * K is row-major 3x3; ext is n_cams x 7 (qx qy qz qw tx ty tz), camera <- body. */
ssba_status ssba_set_cameras(ssba_handle *h, const double K[9], int32_t n_cams,
                             const double *ext_qt);

/* One VertexPose per row; `fixed` may be NULL (= none fixed, as backend.cpp:93-103).
 * Vertex order = vertex-id order (poses before landmarks, backend.cpp:120). */
ssba_status ssba_set_poses(ssba_handle *h, int32_t n_poses, const double *qt,
                           const uint8_t *fixed);

/* One VertexXYZ per row, all setMarginalized(true) (backend.cpp:117-123);
 * `fixed` mirrors setFixed(true) (backend.cpp:125-130), may be NULL. */
ssba_status ssba_set_points(ssba_handle *h, int32_t n_points, const double *xyz,
                            const uint8_t *fixed);

/* One EdgeProjection per row, in addEdge order (backend.cpp:136-168):
 *   pose_idx / point_idx : rows of set_poses / set_points  (setVertex 0 / 1)
 *   cam_idx              : row of ext_qt                   (left/right, :147-155)
 *   uv                   : n_edges x 2 measurement         (setMeasurement)
 *   info                 : n_edges x 3 (xx, xy, yy) or NULL = identity (:161)
 *   huber_delta          : n_edges deltas or NULL => `huber_delta_all` for every edge;
 *                          a delta <= 0 means no robust kernel      (:162-164) */
ssba_status ssba_set_edges(ssba_handle *h, int32_t n_edges, const int32_t *pose_idx,
                           const int32_t *point_idx, const uint8_t *cam_idx,
                           const double *uv, const double *info,
                           const double *huber_delta, double huber_delta_all);

/* ---- optimisation ------------------------------------------------------------------ */

/* SparseOptimizer::initializeOptimization() (sparse_optimizer.cpp:201-272) followed by
 * the structure part of the first solve(): active sets, index mapping, Hessian block
 * pattern, Schur pattern, symbolic factorisation (BlockSolver::buildStructure,
 * g2o/core/block_solver.hpp:102-256; LinearSolverCSparse::computeSymbolicDecomposition,
 * g2o/solvers/csparse/linear_solver_csparse.h:246-308), and the upload to HBM.
 * Called implicitly by ssba_optimize()/ssba_step(0) when the graph changed. */
ssba_status ssba_initialize(ssba_handle *h);

/* SparseOptimizer::optimize(max_iters) (sparse_optimizer.cpp:366-431) with
 * OptimizationAlgorithmLevenberg::solve (levenberg.cpp:58-150). `report` may be NULL. */
ssba_status ssba_optimize(ssba_handle *h, int32_t max_iters, ssba_report *report);

/* One OptimizationAlgorithmLevenberg::solve(iteration) (levenberg.cpp:58-150);
 * iteration == 0 re-initialises lambda like a fresh optimize() call.  This is what
 * the g2o shim (include/ssba_g2o_shim.hpp) forwards solve() to. */
ssba_status ssba_step(ssba_handle *h, int32_t iteration, int32_t *solver_result,
                      ssba_iter_record *record);

/* Restore the estimates given to set_poses/set_points on the device (no host copy);
 * lets a resident graph be optimised repeatedly (bench). */
ssba_status ssba_reset_state(ssba_handle *h);

/* Structure reuse.  When the set_* calls since the last ssba_initialize() changed only VALUES (estimates,
 * measurements, information matrices, Huber widths, camera parameters) and not the topology (vertex counts,
 * fixed flags, edge indices, cameras per edge), ssba_initialize() keeps the resident structure and only
 * uploads the values: what rounds 2..5 of backend.cpp:175-203 need, where g2o re-runs buildStructure and the
 * symbolic factorisation on every optimize() (linear_solver_csparse.h:97-104).  The comparison is exact
 * (memcmp against the arrays of the resident graph).  ssba_drop_structure() forces the next
 * ssba_initialize() to rebuild (benchmarks of the cold path). */
ssba_status ssba_drop_structure(ssba_handle *h);

/* ---- results (backend.cpp:180-244) ------------------------------------------------- */

/* VertexPose::estimate() / VertexXYZ::estimate() (backend.cpp:234,238), same layout
 * and order as the set_* calls. */
ssba_status ssba_get_poses(ssba_handle *h, double *qt_out);
ssba_status ssba_get_points(ssba_handle *h, double *xyz_out);

/* EdgeProjection::error() per edge at the final estimate, n_edges x 2, addEdge order
 * (what ef.first->chi2() reads, backend.cpp:184,209; recomputed at the final state,
 * see SURVEY.md 3.4 note). Inactive edges (both vertices fixed) get 0. */
ssba_status ssba_get_edge_errors(ssba_handle *h, double *err_out);

/* activeChi2() / activeRobustChi2() at the current estimate
 * (sparse_optimizer.cpp:92-116). Either pointer may be NULL. */
ssba_status ssba_chi2(ssba_handle *h, double *plain, double *robust);

/* The inlier/outlier count of backend.cpp:180-197: edges with plain chi2 > threshold
 * are outliers.  Evaluated on the device at the current estimate. */
ssba_status ssba_count_outliers(ssba_handle *h, double chi2_threshold,
                                int64_t *n_outliers, int64_t *n_inliers);

/* The per-edge outlier flags of the culling step (backend.cpp:205-227: `ef.first->chi2() > chi2_th`), evaluated on
 * the device at the current estimate: mask_out[e] = 1 when the plain chi2 of edge e (caller's addEdge order)
 * exceeds the threshold, else 0 (edges between two fixed vertices are never active: 0).  n_edges bytes come back
 * instead of the 16 n_edges of ssba_get_edge_errors.  n_outliers may be NULL. */
ssba_status ssba_get_outlier_mask(ssba_handle *h, double chi2_threshold, uint8_t *mask_out, int64_t *n_outliers);

/* SparseOptimizer::setForceStopFlag (sparse_optimizer.h:183-187): while the flag is raised, optimize()/step() stop
 * like the reference does -- the trial loop of OptimizationAlgorithmLevenberg::solve ends after the running trial
 * (optimization_algorithm_levenberg.cpp:145: `!_optimizer->terminate()`) and optimize() starts no further iteration
 * (sparse_optimizer.cpp:388); raised before the call, optimize() returns 0 iterations.  ssba_request_stop may be
 * called from ANOTHER thread while ssba_optimize runs on the handle (it only posts a 4-byte copy on a side
 * stream); the flag stays up until ssba_clear_stop. */
ssba_status ssba_request_stop(ssba_handle *h);
ssba_status ssba_clear_stop(ssba_handle *h);

/* The optimisation loop of Backend::OptimizeActiveMap (backend.cpp:175-203) with the graph
 * resident on the device between rounds: up to `max_rounds` (5) rounds of
 * initializeOptimization(); optimize(iters_per_round) (10), each followed by the outlier count
 * with `chi2_threshold` (5.891) on the un-robustified chi2; stops as soon as
 * inliers / (inliers + outliers) > `inlier_ratio` (0.7).  lambda is re-initialised in every round,
 * as every optimize() call does.  Outputs may be NULL; `last_report` is the report of the last round. */
ssba_status ssba_optimize_rounds(ssba_handle *h, int32_t max_rounds, int32_t iters_per_round,
                                 double chi2_threshold, double inlier_ratio, int32_t *rounds_done,
                                 int64_t *n_outliers, int64_t *n_inliers, ssba_report *last_report);

/* ---- multi-GPU helpers ------------------------------------------------------------- */

/* Host-only (needs no CUDA device): the landmark shard plan used when world_size > 1.
 * owner_out[point] = rank that owns the landmark and all its edges, -1 if it has no active
 * edge.  Contiguous runs of the landmark order, balanced by edge count. */
ssba_status ssba_plan_shards(int32_t n_poses, const uint8_t *pose_fixed, int32_t n_points,
                             const uint8_t *point_fixed, int32_t n_edges, const int32_t *pose_idx,
                             const int32_t *point_idx, int32_t world_size, int32_t *owner_out);

/* ---- instrumentation --------------------------------------------------------------- */

/* ---- pose-only LM, batched: FrontEnd::EstimateCurrentPose()
 * (src/ssvio/frontend.cpp:184-260) for n_frames frames in one call.  Per frame: one VertexPose, one
 * EdgeProjectionPoseOnly (include/ssvio/g2otypes.hpp:67-110, analytic Jacobian as shipped) per
 * feature with identity information and g2o's default Huber kernel (delta 1), Levenberg with the
 * dense 6x6 solve (BlockSolver_6_3 + LinearSolverDense, frontend.cpp:188-193); `rounds` rounds (the
 * reference: 4) of initializeOptimization(); optimize(iters) (10), each followed by the
 * re-classification of every feature by chi2() > chi2_threshold (5.991; outliers leave the next
 * round, frontend.cpp:243-262), the robust kernel removed after round rounds - 2 (:265-268).
 *   feat_ptr[n_frames + 1]   features of frame f = [feat_ptr[f], feat_ptr[f + 1])
 *   poses_in / poses_out     n_frames x 7 T_cw (qx qy qz qw tx ty tz): initial / optimised
 *   xyz, uv                  per feature: map-point position, measured pixel (K: row-major 3x3)
 *   outlier_out              per feature: features[i]->is_outlier_ after the last round
 *   n_inliers_out            per frame: what EstimateCurrentPose returns
 *   chi2_out                 per frame: activeRobustChi2() after the last optimize (may be NULL)
 * The LM constants are the handle's ssba_options (tau, step scales, max trials).  All arrays are
 * host memory; the call is synchronous.  Runs on the handle's device and stream; the handle's
 * bundle-adjustment problem, if any, is untouched. */
ssba_status ssba_pose_only_optimize(ssba_handle *h, const double K[9], int32_t n_frames,
                                    const int32_t *feat_ptr, const double *poses_in, const double *xyz,
                                    const double *uv, int32_t rounds, int32_t iters, double chi2_threshold,
                                    double *poses_out, uint8_t *outlier_out, int32_t *n_inliers_out,
                                    double *chi2_out);

/* The loop closer's caller of the same pose-only LM: LoopClosing::OptimizeCurrentPose()
 * (src/ssvio/loopclosing.cpp:245-351).  Same vertex, edges, kernel and round loop, preceded by ONE more
 * initializeOptimization(); optimize(iters) with every feature active and no classification (:302-303).
 * Arguments as for ssba_pose_only_optimize (a "frame" here is a loop candidate: the current key-frame's
 * features against the loop key-frame's map points; n_inliers_out = the matches that survive, :337-343). */
ssba_status ssba_pose_only_optimize_loop(ssba_handle *h, const double K[9], int32_t n_frames,
                                    const int32_t *feat_ptr, const double *poses_in, const double *xyz,
                                    const double *uv, int32_t rounds, int32_t iters, double chi2_threshold,
                                    double *poses_out, uint8_t *outlier_out, int32_t *n_inliers_out,
                                    double *chi2_out);

/* ---- pose-graph optimisation: LoopClosing::PoseGraphOptimization()
 * (src/ssvio/loopclosing.cpp:458-532).  One VertexPose per key-frame (setMarginalized(false); `fixed`
 * marks the ones :480-486 fixes), one EdgePoseGraph (include/ssvio/g2otypes.hpp:164-199) per relative-pose
 * constraint: error = log(measurement^-1 * T_v0 * T_v1^-1), identity 6x6 information, no robust kernel,
 * Jacobians numeric as the reference ships them; Levenberg over the block-sparse H (BlockSolver<6,6> +
 * LinearSolverEigen in the reference, here the level-scheduled block solver of the bundle adjustment);
 * initializeOptimization(); optimize(iters) (the reference: 20).
 *   poses_in / poses_out   n_poses x 7 T_cw (qx qy qz qw tx ty tz)
 *   v0, v1, meas           per edge: vertex 0, vertex 1 (pose rows), measurement (7 doubles)
 * report (may be NULL) as for ssba_optimize; chi2_robust = chi2_plain (no kernel).  Host arrays, synchronous;
 * the handle's bundle-adjustment problem, if any, is untouched.  SSBA_ERR_EMPTY when nothing is free. */
ssba_status ssba_pose_graph_optimize(ssba_handle *h, int32_t n_poses, const double *poses_in, const uint8_t *fixed,
                                     int32_t n_edges, const int32_t *v0, const int32_t *v1, const double *meas,
                                     int32_t iters, double *poses_out, ssba_report *report);

/* Turn the per-phase CUDA-event timers on or off after creation (ssba_options::profile sets
 * the initial state).  The timers add event records to the stream, so benchmarks time with
 * them off and profile in a separate pass. */
ssba_status ssba_set_profiling(ssba_handle *h, int32_t on);
ssba_status ssba_profile_get(ssba_handle *h, ssba_profile *out);
ssba_status ssba_profile_reset(ssba_handle *h);

/* Sizes of the resident problem after ssba_initialize(): free poses, free landmarks,
 * active edges, (pose, landmark) pairs, upper-triangular Schur blocks, factor blocks. */
typedef struct ssba_problem_info {
  int32_t n_free_poses, n_free_points, n_active_edges, n_pairs;
  int32_t n_schur_blocks, n_factor_blocks;
  int64_t device_bytes;
  int32_t solve_cluster;   /* CTAs of the thread-block cluster the reduced solve runs on          */
  int32_t peer_exchange;   /* several GPUs: 1 = per-trial exchanges through NVLink peer memory,     */
                           /* 0 = NCCL all-reduce (ranks on different nodes, or CUDA IPC refused)  */
  int32_t solver_kind;     /* 1 = subtree-per-CTA solver (k_tree_solve), 0 = level-scheduled solver  */
  int32_t solver_steps;    /* elimination steps on the critical path of the reduced solve            */
  int32_t solver_top_cols; /* columns of the top part (factored by CTA 0 after the hand-off)          */
  int32_t solver_smem_bytes;
  int64_t n_structure_builds; /* ssba_initialize() calls that built the structure ...               */
  int64_t n_structure_reuses; /* ... and calls that kept it (same topology, values only)             */
} ssba_problem_info;
ssba_status ssba_get_problem_info(ssba_handle *h, ssba_problem_info *out);

int32_t ssba_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SSBA_H */
