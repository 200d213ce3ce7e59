/*
 * ref_harness.h — C entry points of oracle/_ref/libssba_ref.so  (TEST INFRASTRUCTURE).
 *
 * libssba_ref.so is the REFERENCE ITSELF: the vendored g2o core + CSparse solver + CSparse C
 * sources and ssvio's include/ssvio/g2otypes.hpp, compiled from /root/reference by
 * oracle/Makefile, plus ref_harness.cpp which replays the graph construction of
 * src/ssvio/backend.cpp:81-178 over flat arrays (backend.cpp itself needs OpenCV / glog /
 * Pangolin, which this image does not have).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it.  The product (libssba.so) never does.
 */
#ifndef SSBA_REF_HARNESS_H
#define SSBA_REF_HARNESS_H

#include <stdint.h>
#include "ssba.h" /* ssba_report / ssba_iter_record layouts are shared with the product ABI */

#ifdef __cplusplus
extern "C" {
#endif

/* per-phase wall time summed over iterations, from g2o::G2OBatchStatistics
 * (thirdparty/g2o/g2o/core/batch_stats.h:40-77); only when collect_trace != 0 */
typedef struct ssba_ref_stats {
  double t_residuals, t_quadratic_form, t_schur, t_linear_solver, t_linear_solution, t_update;
  double t_initialize; /* initializeOptimization() wall time */
  double t_graph_build; /* new Vertex/Edge + addVertex/addEdge wall time */
  int64_t cholesky_nnz;
  int32_t n_active_edges, n_index_mapping;
} ssba_ref_stats;

/*
 * Build the g2o graph exactly as Backend::OptimizeActiveMap() does and run
 * initializeOptimization(); optimize(max_iters).  Argument layout = include/ssba.h.
 *
 *   jacobian_mode  0 = analytic (corrected closed form, subclass override of linearizeOplus)
 *                  1 = numeric central differences = the reference AS SHIPPED
 *   collect_trace  1 = record chi2/lambda/trials per iteration (adds one computeActiveErrors per
 *                      iteration, so do not use for timing), 0 = timing run
 *   rounds         number of initializeOptimization()+optimize() rounds of backend.cpp:175-203
 *                  (<=0 or 1: one round, no inlier-ratio rule); with rounds > 1 the loop stops
 *                  early when the inlier ratio exceeds 0.7 like the reference
 * Outputs may be NULL.  report->seconds_total = wall time of the optimize() calls only,
 * report->seconds_setup = graph construction + initializeOptimization().
 * Returns 0 on success.
 */
int ssba_ref_optimize(const double K[9], int32_t n_cams, const double *ext_qt,
                      int32_t n_poses, const double *poses_qt, const uint8_t *pose_fixed,
                      int32_t n_points, const double *points, const uint8_t *point_fixed,
                      int32_t n_edges, const int32_t *pose_idx, const int32_t *point_idx,
                      const uint8_t *cam_idx, const double *uv, double huber_delta,
                      int32_t max_iters, int32_t jacobian_mode, int32_t collect_trace,
                      int32_t rounds, double outlier_chi2_threshold,
                      double *poses_out, double *points_out, double *edge_err_out,
                      ssba_report *report, ssba_ref_stats *stats, int32_t *rounds_done,
                      int64_t *n_outliers);

/* Extras of the NEXT ssba_ref_optimize call (consumed by it): _userLambdaInit and _maxTrialsAfterFailure of
 * OptimizationAlgorithmLevenberg (thirdparty/g2o/g2o/core/optimization_algorithm_levenberg.cpp:44-56; 0 keeps
 * the default) and per-edge information matrices (n_edges x 3: xx, xy, yy; NULL = identity as backend.cpp:161).
 * Used to drive the reference into failed factorisations (csparse_extension.cpp:115). */
void ssba_ref_set_extras(double user_lambda_init, int32_t max_trials_after_failure, const double *edge_info_xx_xy_yy);

/*
 * Pose-only LM of FrontEnd::EstimateCurrentPose() (src/ssvio/frontend.cpp:184-260), frame by
 * frame: one VertexPose, one EdgeProjectionPoseOnly (include/ssvio/g2otypes.hpp:67-110, analytic
 * Jacobian as shipped) per feature with identity information and a default RobustKernelHuber
 * (delta = 1), BlockSolver_6_3 + LinearSolverDense + Levenberg; `rounds` rounds (4 in the
 * reference) of initializeOptimization(); optimize(iters) (10), after each of which every
 * feature is re-classified by chi2() > chi2_threshold (5.991; outliers get level 1 and their error
 * recomputed before the test), and the robust kernel is removed at the end of round rounds - 2.
 *
 *   feat_ptr[n_frames + 1]  features of frame f = [feat_ptr[f], feat_ptr[f + 1])
 *   poses_qt                n_frames x 7 initial T_cw (qx qy qz qw tx ty tz)
 *   xyz, uv                 per feature: map-point position, measured pixel
 * Outputs (may be NULL): optimised poses, per-feature outlier flag, per-frame inlier count
 * (what EstimateCurrentPose returns), per-frame activeRobustChi2 after the last round's optimize.
 * Returns 0 on success.
 */
int ssba_ref_pose_only(const double K[9], int32_t n_frames, const int32_t *feat_ptr,
                       const double *poses_qt, const double *xyz, const double *uv,
                       int32_t rounds, int32_t iters, double chi2_threshold,
                       double *poses_out, uint8_t *outlier_out, int32_t *n_inliers_out,
                       double *chi2_out);

/* The same with `pre_rounds` unclassified initializeOptimization(); optimize(iters) calls before the rounds:
 * pre_rounds = 1 is LoopClosing::OptimizeCurrentPose() (src/ssvio/loopclosing.cpp:245-351, the extra call at :302-303). */
int ssba_ref_pose_only_ex(const double K[9], int32_t n_frames, const int32_t *feat_ptr,
                          const double *poses_qt, const double *xyz, const double *uv,
                          int32_t rounds, int32_t iters, double chi2_threshold, int32_t pre_rounds,
                          double *poses_out, uint8_t *outlier_out, int32_t *n_inliers_out,
                          double *chi2_out);

/*
 * Pose-graph optimisation of LoopClosing::PoseGraphOptimization() (src/ssvio/loopclosing.cpp:458-594):
 * one VertexPose per key-frame (setMarginalized(false), some fixed), one EdgePoseGraph
 * (include/ssvio/g2otypes.hpp:164-199: error = log(measurement^-1 * v0 * v1^-1), numeric Jacobians as
 * shipped) with identity 6x6 information and no robust kernel per (vertex 0, vertex 1, measurement),
 * BlockSolver<6,6> + LinearSolverEigen + Levenberg, initializeOptimization(); optimize(iters) (20).
 *   poses_qt n_poses x 7 (T_cw), fixed n_poses, edges: v0[e], v1[e], meas_qt n_edges x 7
 * Outputs may be NULL.  report as for ssba_ref_optimize (chi2 = activeRobustChi2 = plain chi2 here).
 */
int ssba_ref_pose_graph(int32_t n_poses, const double *poses_qt, const uint8_t *fixed, int32_t n_edges,
                        const int32_t *v0, const int32_t *v1, const double *meas_qt, int32_t iters,
                        double *poses_out, ssba_report *report);

#ifdef __cplusplus
}
#endif
#endif
