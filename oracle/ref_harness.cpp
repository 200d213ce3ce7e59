// ref_harness.cpp — drives the UNMODIFIED reference (g2o + ssvio g2otypes.hpp, compiled from
// /root/reference by oracle/Makefile) over flat arrays.  TEST INFRASTRUCTURE, see ref_harness.h.
//
// Follows src/ssvio/backend.cpp:81-203 step by step: solver stack (:81-86), VertexPose per
// key-frame (:93-103), VertexXYZ per map-point with setMarginalized(true) and the fixed rule
// (:113-133), one EdgeProjection per observation with identity information and a Huber kernel
// (:136-168), then rounds of initializeOptimization(); optimize(N) with the inlier-ratio rule
// (:175-203).  The only liberties: map/key-frame containers are flat arrays, vertex ids are
// (pose row) and (n_poses + point row), and the final edge errors are recomputed at the final
// estimate before they are read (SURVEY.md 3.4 note).

#include <chrono>
#include <cstring>
#include <memory>
#include <vector>

#include "ssvio/g2otypes.hpp"
#include "g2o/solvers/dense/linear_solver_dense.h"
#include "g2o/solvers/eigen/linear_solver_eigen.h"

#include "ref_harness.h"

namespace {

using Clock = std::chrono::steady_clock;
inline double seconds_since(Clock::time_point t0) {
  return std::chrono::duration<double>(Clock::now() - t0).count();
}

// Corrected closed-form Jacobian: the commented-out override in
// include/ssvio/g2otypes.hpp:133-153 is wrong for the right camera; this one is exact.
class EdgeProjectionAnalytic : public ssvio::EdgeProjection {
 public:
  EIGEN_MAKE_ALIGNED_OPERATOR_NEW;
  EdgeProjectionAnalytic(const Eigen::Matrix3d &K, const Sophus::SE3d &ext)
      : ssvio::EdgeProjection(K, ext), K_(K), ext_(ext) {}

  void linearizeOplus() override {
    const auto *v0 = static_cast<const ssvio::VertexPose *>(_vertices[0]);
    const auto *v1 = static_cast<const ssvio::VertexXYZ *>(_vertices[1]);
    const Sophus::SE3d T = v0->estimate();
    const Eigen::Vector3d pb = T * v1->estimate();  // body frame
    const Eigen::Vector3d pc = ext_ * pb;           // camera frame
    const double fx = K_(0, 0), fy = K_(1, 1);
    const double X = pc[0], Y = pc[1], Z = pc[2];
    const double zi = 1.0 / Z, zi2 = zi * zi;
    Eigen::Matrix<double, 2, 3> Jpi;
    Jpi << -fx * zi, 0.0, fx * X * zi2, 0.0, -fy * zi, fy * Y * zi2;
    const Eigen::Matrix<double, 2, 3> JR = Jpi * ext_.rotationMatrix();
    _jacobianOplusXi.block<2, 3>(0, 0) = JR;
    _jacobianOplusXi.block<2, 3>(0, 3) = -JR * Sophus::SO3d::hat(pb);
    _jacobianOplusXj = JR * T.rotationMatrix();
  }

 private:
  Eigen::Matrix3d K_;
  Sophus::SE3d ext_;
};

Sophus::SE3d se3_from_qt(const double *qt) {
  // memory order of the C ABI: qx qy qz qw tx ty tz. Eigen::Quaterniond ctor takes (w,x,y,z).
  Eigen::Quaterniond q(qt[3], qt[0], qt[1], qt[2]);
  return Sophus::SE3d(q, Eigen::Vector3d(qt[4], qt[5], qt[6]));
}

void qt_from_se3(const Sophus::SE3d &T, double *qt) {
  const Eigen::Quaterniond &q = T.unit_quaternion();
  qt[0] = q.x(); qt[1] = q.y(); qt[2] = q.z(); qt[3] = q.w();
  qt[4] = T.translation()[0]; qt[5] = T.translation()[1]; qt[6] = T.translation()[2];
}

struct TraceAction : public g2o::HyperGraphAction {
  g2o::SparseOptimizer *opt = nullptr;
  g2o::OptimizationAlgorithmLevenberg *lm = nullptr;
  ssba_report *report = nullptr;
  g2o::HyperGraphAction *operator()(const g2o::HyperGraph *, Parameters *p = 0) override {
    auto *pi = dynamic_cast<ParametersIteration *>(p);
    if (!pi || pi->iteration < 0 || !report) return this;
    if (report->n_records >= SSBA_MAX_ITER_RECORDS) return this;
    ssba_iter_record &r = report->iters[report->n_records++];
    opt->computeActiveErrors();
    r.chi2 = opt->activeRobustChi2();
    r.lambda = lm->currentLambda();
    r.trials = lm->levenbergIteration();
    r.result = SSBA_SOLVER_OK;  // g2o does not expose the per-iteration result; see iterations
    return this;
  }
};

// LinearSolverCSparse that counts its failed factorisations (csparse_extension.cpp:115 -> solve() == false):
// g2o itself only turns them into a rejected trial (optimization_algorithm_levenberg.cpp:120-121)
int g_cholesky_failures = 0;
template <class M>
class CountingCSparse : public g2o::LinearSolverCSparse<M> {
 public:
  bool solve(const g2o::SparseBlockMatrix<M> &A, number_t *x, number_t *b) override {
    const bool ok = g2o::LinearSolverCSparse<M>::solve(A, x, b);
    if (!ok) ++g_cholesky_failures;
    return ok;
  }
};

// extras of the NEXT ssba_ref_optimize call (consumed by it): LM options of
// OptimizationAlgorithmLevenberg (optimization_algorithm_levenberg.cpp:44-56) and per-edge information
// matrices (EdgeProjection::setInformation, backend.cpp:161 uses the identity)
double g_user_lambda_init = 0.0;
int g_max_trials = 0;
const double *g_edge_info = nullptr;

}  // namespace

extern "C" void ssba_ref_set_extras(double user_lambda_init, int32_t max_trials_after_failure, const double *edge_info_xx_xy_yy) {
  g_user_lambda_init = user_lambda_init;
  g_max_trials = max_trials_after_failure;
  g_edge_info = edge_info_xx_xy_yy;
}

extern "C" int ssba_ref_optimize(
    const double K[9], int32_t n_cams, const double *ext_qt, int32_t n_poses,
    const double *poses_qt, const uint8_t *pose_fixed, int32_t n_points, const double *points,
    const uint8_t *point_fixed, int32_t n_edges, const int32_t *pose_idx,
    const int32_t *point_idx, const uint8_t *cam_idx, const double *uv, double huber_delta,
    int32_t max_iters, int32_t jacobian_mode, int32_t collect_trace, int32_t rounds,
    double outlier_chi2_threshold, double *poses_out, double *points_out, double *edge_err_out,
    ssba_report *report, ssba_ref_stats *stats, int32_t *rounds_done, int64_t *n_outliers) {
  if (!K || !ext_qt || !poses_qt || !points || (n_edges > 0 && (!pose_idx || !point_idx || !uv)))
    return 1;
  if (report) std::memset(report, 0, sizeof(*report));
  if (stats) std::memset(stats, 0, sizeof(*stats));

  auto t_build0 = Clock::now();

  // backend.cpp:81-86
  typedef g2o::BlockSolver_6_3 BlockSolverType;
  typedef CountingCSparse<BlockSolverType::PoseMatrixType> LinearSolverType;  // LinearSolverCSparse + a failure counter
  g_cholesky_failures = 0;
  auto *solver = new g2o::OptimizationAlgorithmLevenberg(
      g2o::make_unique<BlockSolverType>(g2o::make_unique<LinearSolverType>()));
  g2o::SparseOptimizer optimizer;
  optimizer.setAlgorithm(solver);
  if (g_user_lambda_init > 0) solver->setUserLambdaInit(g_user_lambda_init);
  if (g_max_trials > 0) solver->setMaxTrialsAfterFailure(g_max_trials);
  solver->setWriteDebug(false);  // a failed factorisation must not drop debug.txt into the working directory
  const double *edge_info = g_edge_info;
  g_user_lambda_init = 0.0; g_max_trials = 0; g_edge_info = nullptr;

  Eigen::Matrix3d cam_K;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) cam_K(r, c) = K[3 * r + c];
  std::vector<Sophus::SE3d, Eigen::aligned_allocator<Sophus::SE3d>> ext(n_cams);
  for (int c = 0; c < n_cams; ++c) ext[c] = se3_from_qt(ext_qt + 7 * c);

  // backend.cpp:93-103
  std::vector<ssvio::VertexPose *> vpose(n_poses);
  for (int i = 0; i < n_poses; ++i) {
    auto *v = new ssvio::VertexPose();
    v->setId(i);
    v->setEstimate(se3_from_qt(poses_qt + 7 * i));
    if (pose_fixed && pose_fixed[i]) v->setFixed(true);  // only BASELINE cfg5 fixes KF0
    optimizer.addVertex(v);
    vpose[i] = v;
  }
  // backend.cpp:113-133
  std::vector<ssvio::VertexXYZ *> vpoint(n_points);
  for (int j = 0; j < n_points; ++j) {
    auto *v = new ssvio::VertexXYZ;
    v->setEstimate(Eigen::Vector3d(points[3 * j], points[3 * j + 1], points[3 * j + 2]));
    v->setId(n_poses + j);
    v->setMarginalized(true);
    if (point_fixed && point_fixed[j]) v->setFixed(true);
    optimizer.addVertex(v);
    vpoint[j] = v;
  }
  // backend.cpp:136-168
  std::vector<ssvio::EdgeProjection *> edges(n_edges);
  for (int e = 0; e < n_edges; ++e) {
    const int cam = cam_idx ? cam_idx[e] : 0;
    ssvio::EdgeProjection *edge = nullptr;
    if (jacobian_mode == SSBA_JACOBIAN_ANALYTIC)
      edge = new EdgeProjectionAnalytic(cam_K, ext[cam]);
    else
      edge = new ssvio::EdgeProjection(cam_K, ext[cam]);
    edge->setId(e + 1);
    edge->setVertex(0, vpose[pose_idx[e]]);
    edge->setVertex(1, vpoint[point_idx[e]]);
    edge->setMeasurement(Eigen::Vector2d(uv[2 * e], uv[2 * e + 1]));
    if (edge_info) {
      Eigen::Matrix2d om;
      om << edge_info[3 * e], edge_info[3 * e + 1], edge_info[3 * e + 1], edge_info[3 * e + 2];
      edge->setInformation(om);
    } else {
      edge->setInformation(Eigen::Matrix2d::Identity());
    }
    if (huber_delta > 0) {
      auto *rk = new g2o::RobustKernelHuber();
      rk->setDelta(huber_delta);
      edge->setRobustKernel(rk);
    }
    optimizer.addEdge(edge);
    edges[e] = edge;
  }
  const double t_build = seconds_since(t_build0);

  TraceAction trace;
  trace.opt = &optimizer;
  trace.lm = solver;
  trace.report = report;
  if (collect_trace) {
    optimizer.addPostIterationAction(&trace);
    optimizer.setComputeBatchStatistics(true);
  }

  // backend.cpp:175-203
  const int max_rounds = rounds > 1 ? rounds : 1;
  double t_init = 0, t_opt = 0;
  int its = 0, round = 0;
  int64_t cnt_outlier = 0, cnt_inlier = 0;
  bool first = true;
  while (round < max_rounds) {
    auto t0 = Clock::now();
    bool ok = optimizer.initializeOptimization();
    t_init += seconds_since(t0);
    if (first && report) {
      if (ok) {
        optimizer.computeActiveErrors();
        report->chi2_initial = optimizer.activeRobustChi2();
      }
      first = false;
    }
    auto t1 = Clock::now();
    its = optimizer.optimize(max_iters);
    t_opt += seconds_since(t1);
    if (collect_trace && stats) {
      for (const auto &s : optimizer.batchStatistics()) {
        stats->t_residuals += s.timeResiduals;
        stats->t_quadratic_form += s.timeQuadraticForm;
        stats->t_schur += s.timeSchurComplement;
        stats->t_linear_solver += s.timeLinearSolver;
        stats->t_linear_solution += s.timeLinearSolution;
        stats->t_update += s.timeUpdate;
        stats->cholesky_nnz = (int64_t)s.choleskyNNZ;
      }
    }
    ++round;
    if (its < 0) break;
    // the reference reads possibly stale _error here (no recompute after a rejected last
    // trial); the harness recomputes first so the count is a function of the estimate.
    optimizer.computeActiveErrors();
    cnt_outlier = cnt_inlier = 0;
    for (auto *e : edges) {
      if (e->chi2() > outlier_chi2_threshold) ++cnt_outlier; else ++cnt_inlier;
    }
    if (max_rounds == 1) break;
    const double inlier_ratio = cnt_inlier / double(cnt_inlier + cnt_outlier);
    if (inlier_ratio > 0.7) break;
  }
  if (rounds_done) *rounds_done = round;
  if (n_outliers) *n_outliers = cnt_outlier;

  if (its >= 0) optimizer.computeActiveErrors();
  if (report) {
    report->iterations = its;
    report->last_result = its > 0 ? SSBA_SOLVER_OK : SSBA_SOLVER_FAIL;
    report->chi2_robust = its >= 0 ? optimizer.activeRobustChi2() : 0.0;
    report->chi2_plain = its >= 0 ? optimizer.activeChi2() : 0.0;
    report->lambda = solver->currentLambda();
    report->cholesky_failures = g_cholesky_failures;
    report->seconds_total = t_opt;
    report->seconds_setup = t_build + t_init;
  }
  if (stats) {
    stats->t_initialize = t_init;
    stats->t_graph_build = t_build;
    stats->n_active_edges = (int32_t)optimizer.activeEdges().size();
    stats->n_index_mapping = (int32_t)optimizer.indexMapping().size();
  }
  if (poses_out)
    for (int i = 0; i < n_poses; ++i) qt_from_se3(vpose[i]->estimate(), poses_out + 7 * i);
  if (points_out)
    for (int j = 0; j < n_points; ++j) {
      const Eigen::Vector3d &p = vpoint[j]->estimate();
      points_out[3 * j] = p[0]; points_out[3 * j + 1] = p[1]; points_out[3 * j + 2] = p[2];
    }
  if (edge_err_out) {
    // active edges hold a fresh error; inactive ones (both vertices fixed) report 0
    std::vector<char> active(n_edges, 0);
    for (auto *ae : optimizer.activeEdges()) active[ae->id() - 1] = 1;
    for (int e = 0; e < n_edges; ++e) {
      edge_err_out[2 * e] = active[e] ? edges[e]->error()[0] : 0.0;
      edge_err_out[2 * e + 1] = active[e] ? edges[e]->error()[1] : 0.0;
    }
  }
  if (collect_trace) optimizer.removePostIterationAction(&trace);
  return 0;
}


// ---- FrontEnd::EstimateCurrentPose (src/ssvio/frontend.cpp:184-260) over flat arrays, one frame
// at a time exactly like the front-end thread does it: the solver stack of :188-193, the vertex
// of :196-203, the edges of :206-231 (identity information, default-delta Huber), the round loop
// of :235-270.  features[i]->is_outlier_ lives in `outlier`.
// pre_rounds = 1: LoopClosing::OptimizeCurrentPose (src/ssvio/loopclosing.cpp:245-351), which differs from the
// front-end's loop by one unclassified initializeOptimization(); optimize(10) before the rounds (:302-303)
extern "C" int ssba_ref_pose_only_ex(const double K9[9], int32_t n_frames, const int32_t *feat_ptr,
                                     const double *poses_qt, const double *xyz, const double *uv,
                                     int32_t rounds, int32_t iters, double chi2_threshold, int32_t pre_rounds,
                                     double *poses_out, uint8_t *outlier_out, int32_t *n_inliers_out,
                                     double *chi2_out) {
  Eigen::Matrix3d K;
  K << K9[0], K9[1], K9[2], K9[3], K9[4], K9[5], K9[6], K9[7], K9[8];
  for (int f = 0; f < n_frames; ++f) {
    typedef g2o::BlockSolver_6_3 BlockSolverType;
    typedef g2o::LinearSolverDense<BlockSolverType::PoseMatrixType> LinearSolverType;
    auto solver = new g2o::OptimizationAlgorithmLevenberg(
        g2o::make_unique<BlockSolverType>(g2o::make_unique<LinearSolverType>()));
    g2o::SparseOptimizer optimizer;
    optimizer.setAlgorithm(solver);
    ssvio::VertexPose *vertex_pose = new ssvio::VertexPose();
    vertex_pose->setId(0);
    vertex_pose->setEstimate(se3_from_qt(poses_qt + 7 * f));
    optimizer.addVertex(vertex_pose);
    const int e0 = feat_ptr[f], n = feat_ptr[f + 1] - e0;
    std::vector<ssvio::EdgeProjectionPoseOnly *> edges;
    std::vector<uint8_t> outlier(n, 0);
    edges.reserve(n);
    int index = 1;
    for (int i = 0; i < n; ++i) {
      auto *edge = new ssvio::EdgeProjectionPoseOnly(
          Eigen::Vector3d(xyz[3 * (e0 + i)], xyz[3 * (e0 + i) + 1], xyz[3 * (e0 + i) + 2]), K);
      edge->setId(index);
      edge->setVertex(0, vertex_pose);
      edge->setMeasurement(Eigen::Vector2d(uv[2 * (e0 + i)], uv[2 * (e0 + i) + 1]));
      edge->setInformation(Eigen::Matrix2d::Identity());
      edge->setRobustKernel(new g2o::RobustKernelHuber);
      edges.emplace_back(edge);
      optimizer.addEdge(edge);
      index++;
    }
    int cnt_outliers = 0;
    double chi_last = 0.0;
    for (int pre = 0; pre < pre_rounds; ++pre) {  // loopclosing.cpp:302-303
      optimizer.initializeOptimization();
      optimizer.optimize(iters);
      chi_last = optimizer.activeRobustChi2();
    }
    for (int iteration = 0; iteration < rounds; iteration++) {
      optimizer.initializeOptimization();
      optimizer.optimize(iters);
      chi_last = optimizer.activeRobustChi2();
      cnt_outliers = 0;
      for (int i = 0; i < n; i++) {
        auto e = edges[i];
        if (outlier[i]) e->computeError();
        if (e->chi2() > chi2_threshold) {
          outlier[i] = 1;
          e->setLevel(1);
          cnt_outliers++;
        } else {
          outlier[i] = 0;
          e->setLevel(0);
        }
        if (iteration == rounds - 2) e->setRobustKernel(nullptr);
      }
    }
    if (poses_out) qt_from_se3(vertex_pose->estimate(), poses_out + 7 * f);
    if (outlier_out) std::memcpy(outlier_out + e0, outlier.data(), n);
    if (n_inliers_out) n_inliers_out[f] = n - cnt_outliers;
    if (chi2_out) chi2_out[f] = chi_last;
  }
  return 0;
}


extern "C" int ssba_ref_pose_only(const double K9[9], int32_t n_frames, const int32_t *feat_ptr,
                                  const double *poses_qt, const double *xyz, const double *uv,
                                  int32_t rounds, int32_t iters, double chi2_threshold,
                                  double *poses_out, uint8_t *outlier_out, int32_t *n_inliers_out,
                                  double *chi2_out) {
  return ssba_ref_pose_only_ex(K9, n_frames, feat_ptr, poses_qt, xyz, uv, rounds, iters, chi2_threshold, 0, poses_out,
                               outlier_out, n_inliers_out, chi2_out);
}


// ---- LoopClosing::PoseGraphOptimization (src/ssvio/loopclosing.cpp:458-532) over flat arrays: the solver
// stack of :460-465, one vertex per key-frame (:471-489), one EdgePoseGraph per relative-pose constraint
// with identity information (:497-527), initializeOptimization(); optimize(iters) (:531-532).
extern "C" int ssba_ref_pose_graph(int32_t n_poses, const double *poses_qt, const uint8_t *fixed, int32_t n_edges,
                                   const int32_t *v0, const int32_t *v1, const double *meas_qt, int32_t iters,
                                   double *poses_out, ssba_report *report) {
  typedef g2o::BlockSolver<g2o::BlockSolverTraits<6, 6>> BlockSolverType;
  typedef g2o::LinearSolverEigen<BlockSolverType::PoseMatrixType> LinearSolverType;
  auto solver = new g2o::OptimizationAlgorithmLevenberg(
      g2o::make_unique<BlockSolverType>(g2o::make_unique<LinearSolverType>()));
  g2o::SparseOptimizer optimizer;
  optimizer.setAlgorithm(solver);
  std::vector<ssvio::VertexPose *> vertices(n_poses);
  for (int i = 0; i < n_poses; ++i) {
    auto *vertex_pose = new ssvio::VertexPose();
    vertex_pose->setId(i);
    vertex_pose->setEstimate(se3_from_qt(poses_qt + 7 * i));
    vertex_pose->setMarginalized(false);
    if (fixed && fixed[i]) vertex_pose->setFixed(true);
    optimizer.addVertex(vertex_pose);
    vertices[i] = vertex_pose;
  }
  for (int e = 0; e < n_edges; ++e) {
    auto *edge = new ssvio::EdgePoseGraph();
    edge->setId(e);
    edge->setVertex(0, vertices[v0[e]]);
    edge->setVertex(1, vertices[v1[e]]);
    edge->setMeasurement(se3_from_qt(meas_qt + 7 * e));
    edge->setInformation(Eigen::Matrix<double, 6, 6>::Identity());
    optimizer.addEdge(edge);
  }
  if (report) std::memset(report, 0, sizeof(*report));
  TraceAction trace;
  trace.opt = &optimizer; trace.lm = solver; trace.report = report;
  if (report) optimizer.addPostIterationAction(&trace);
  optimizer.initializeOptimization();
  if (report) { optimizer.computeActiveErrors(); report->chi2_initial = optimizer.activeRobustChi2(); }
  const auto t0 = Clock::now();
  const int its = optimizer.optimize(iters);
  if (report) {
    report->seconds_total = seconds_since(t0);
    report->iterations = its;
    optimizer.computeActiveErrors();
    report->chi2_robust = optimizer.activeRobustChi2();
    report->chi2_plain = optimizer.activeChi2();
    report->lambda = solver->currentLambda();
  }
  if (poses_out)
    for (int i = 0; i < n_poses; ++i) qt_from_se3(vertices[i]->estimate(), poses_out + 7 * i);
  return 0;
}
