// ssba_g2o_shim.hpp — header-only g2o plug-in that puts libssba (B200, sm_100a) behind the
// reference's own optimisation-algorithm interface, so that ssvio's Backend::OptimizeActiveMap()
// (src/ssvio/backend.cpp:78-245) keeps its vertex / edge API and changes by one typedef:
//
//   - auto solver = new g2o::OptimizationAlgorithmLevenberg(
//   -     g2o::make_unique<BlockSolverType>(g2o::make_unique<LinearSolverType>()));   // :83-84
//   + auto solver = new ssba::OptimizationAlgorithmLevenbergCuda<>();
//     g2o::SparseOptimizer optimizer;
//     optimizer.setAlgorithm(solver);                                                  // :85-86
//
// It implements g2o::OptimizationAlgorithm (thirdparty/g2o/g2o/core/optimization_algorithm.h:50-81):
//   init()   flattens the ACTIVE graph that SparseOptimizer::initializeOptimization() prepared
//            (activeVertices()/activeEdges(), sparse_optimizer.cpp:201-272) into the flat arrays
//            of include/ssba.h and uploads it;
//   solve(i) runs one outer Levenberg iteration on the GPU (ssba_step) = what
//            OptimizationAlgorithmLevenberg::solve does (optimization_algorithm_levenberg.cpp:58-150)
//            and returns the same SolverResult; estimates and per-edge errors are written back
//            into the g2o vertices / edges so that v->estimate() (backend.cpp:234,238) and
//            e->chi2() (backend.cpp:184,209) keep working unchanged.
//
// The only thing the shim cannot read from an unmodified ssvio::EdgeProjection is its private
// (K, cam_ext) pair (include/ssvio/g2otypes.hpp:159-161).  `EdgeAccess` supplies them; the
// default expects the two const accessors shown in INTEGRATION.md (a 2-line, ABI-neutral patch
// to the header-only class).
//
// Needs: g2o core headers, ssvio/g2otypes.hpp, ssba.h; link with -lssba.
#pragma once

#include <cstring>
#include <iostream>
#include <unordered_map>
#include <vector>

#include <g2o/core/optimization_algorithm.h>
#include <g2o/core/robust_kernel_impl.h>
#include <g2o/core/sparse_optimizer.h>

#include "ssba.h"

namespace ssba {

// Default access to the camera of an edge: requires
//   const Eigen::Matrix3d &K() const { return _K; }
//   const Sophus::SE3d &cam_ext() const { return _cam_ext; }
// on ssvio::EdgeProjection.
struct AccessorEdgeAccess {
  template <class Edge> static const Eigen::Matrix3d &K(const Edge *e) { return e->K(); }
  template <class Edge> static const Sophus::SE3d &ext(const Edge *e) { return e->cam_ext(); }
};

template <class VertexPoseT, class VertexXYZT, class EdgeProjectionT, class EdgeAccess = AccessorEdgeAccess>
class OptimizationAlgorithmLevenbergCudaT : public g2o::OptimizationAlgorithm {
 public:
  explicit OptimizationAlgorithmLevenbergCudaT(const ssba_options *opt = nullptr) {
    if (opt) opt_ = *opt; else ssba_default_options(&opt_);
  }
  ~OptimizationAlgorithmLevenbergCudaT() override { ssba_destroy(h_); }

  // OptimizationAlgorithmWithHessian::init (optimization_algorithm_with_hessian.cpp:48-73)
  bool init(bool /*online*/ = false) override {
    if (!_optimizer) return false;
    if (!h_ && ssba_create(&opt_, &h_) != SSBA_OK) {
      std::cerr << "ssba: " << ssba_last_error(nullptr) << std::endl;
      return false;
    }
    poses_.clear(); points_.clear(); edges_.clear();
    std::unordered_map<const g2o::HyperGraph::Vertex *, int32_t> row;
    std::vector<double> pose_qt, xyz;
    std::vector<uint8_t> pose_fixed, point_fixed;
    for (auto *v : _optimizer->activeVertices()) {  // sorted by id (sparse_optimizer.cpp:493-498)
      if (auto *vp = dynamic_cast<VertexPoseT *>(v)) {
        row[v] = (int32_t)poses_.size();
        poses_.push_back(vp);
        const auto &T = vp->estimate();
        const auto &q = T.unit_quaternion();
        const double qt[7] = {q.x(), q.y(), q.z(), q.w(), T.translation()[0], T.translation()[1], T.translation()[2]};
        pose_qt.insert(pose_qt.end(), qt, qt + 7);
        pose_fixed.push_back(vp->fixed());
      } else if (auto *vl = dynamic_cast<VertexXYZT *>(v)) {
        row[v] = (int32_t)points_.size();
        points_.push_back(vl);
        const auto &p = vl->estimate();
        xyz.insert(xyz.end(), {p[0], p[1], p[2]});
        point_fixed.push_back(vl->fixed());
      } else {
        std::cerr << "ssba: unsupported vertex type in the active graph" << std::endl;
        return false;
      }
    }
    std::vector<int32_t> pidx, lidx;
    std::vector<uint8_t> cam;
    std::vector<double> uv, info, delta, ext_qt;
    double K[9] = {0};
    bool have_K = false;
    for (auto *e : _optimizer->activeEdges()) {  // internalId = addEdge order
      auto *ep = dynamic_cast<EdgeProjectionT *>(e);
      if (!ep) { std::cerr << "ssba: unsupported edge type in the active graph" << std::endl; return false; }
      edges_.push_back(ep);
      pidx.push_back(row.at(ep->vertex(0)));
      lidx.push_back(row.at(ep->vertex(1)));
      const Eigen::Matrix3d &Ke = EdgeAccess::K(ep);
      if (!have_K) { for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) K[3 * r + c] = Ke(r, c); have_K = true; }
      const auto &X = EdgeAccess::ext(ep);
      const auto &q = X.unit_quaternion();
      const double qt[7] = {q.x(), q.y(), q.z(), q.w(), X.translation()[0], X.translation()[1], X.translation()[2]};
      int ci = -1;
      for (size_t c = 0; c < ext_qt.size() / 7; ++c)
        if (std::memcmp(&ext_qt[7 * c], qt, sizeof(qt)) == 0) { ci = (int)c; break; }
      if (ci < 0) {
        if (ext_qt.size() / 7 >= SSBA_MAX_CAMERAS) { std::cerr << "ssba: too many distinct camera extrinsics" << std::endl; return false; }
        ci = (int)(ext_qt.size() / 7);
        ext_qt.insert(ext_qt.end(), qt, qt + 7);
      }
      cam.push_back((uint8_t)ci);
      uv.insert(uv.end(), {ep->measurement()[0], ep->measurement()[1]});
      const auto &I = ep->information();
      info.insert(info.end(), {I(0, 0), I(0, 1), I(1, 1)});
      double d = 0.0;
      if (auto *rk = ep->robustKernel()) {
        auto *hub = dynamic_cast<g2o::RobustKernelHuber *>(rk);
        if (!hub) { std::cerr << "ssba: only RobustKernelHuber is supported" << std::endl; return false; }
        d = hub->delta();
      }
      delta.push_back(d);
    }
    if (edges_.empty()) return false;
    bool ok = ssba_set_cameras(h_, K, (int32_t)(ext_qt.size() / 7), ext_qt.data()) == SSBA_OK &&
              ssba_set_poses(h_, (int32_t)poses_.size(), pose_qt.data(), pose_fixed.data()) == SSBA_OK &&
              ssba_set_points(h_, (int32_t)points_.size(), xyz.data(), point_fixed.data()) == SSBA_OK &&
              ssba_set_edges(h_, (int32_t)edges_.size(), pidx.data(), lidx.data(), cam.data(), uv.data(),
                             info.data(), delta.data(), 0.0) == SSBA_OK &&
              ssba_initialize(h_) == SSBA_OK;
    if (!ok) std::cerr << "ssba: " << ssba_last_error(h_) << std::endl;
    return ok;
  }

  // OptimizationAlgorithmLevenberg::solve (optimization_algorithm_levenberg.cpp:58-150)
  SolverResult solve(int iteration, bool /*online*/ = false) override {
    int32_t res = SSBA_SOLVER_FAIL;
    if (ssba_step(h_, iteration, &res, &last_) != SSBA_OK) {
      std::cerr << "ssba: " << ssba_last_error(h_) << std::endl;
      return Fail;
    }
    if (write_back_every_iteration_ && !writeBack()) return Fail;
    return res == SSBA_SOLVER_OK ? OK : res == SSBA_SOLVER_TERMINATE ? Terminate : Fail;
  }

  // estimates -> g2o vertices, errors -> g2o edges (what push/pop/update and computeActiveErrors
  // leave behind in the reference)
  bool writeBack() {
    std::vector<double> qt(7 * poses_.size()), xyz(3 * points_.size()), err(2 * edges_.size());
    if (ssba_get_poses(h_, qt.data()) != SSBA_OK || ssba_get_points(h_, xyz.data()) != SSBA_OK ||
        ssba_get_edge_errors(h_, err.data()) != SSBA_OK) {
      std::cerr << "ssba: " << ssba_last_error(h_) << std::endl;
      return false;
    }
    for (size_t i = 0; i < poses_.size(); ++i) {
      if (poses_[i]->fixed()) continue;
      const double *p = &qt[7 * i];
      poses_[i]->setEstimate(Sophus::SE3d(Eigen::Quaterniond(p[3], p[0], p[1], p[2]), Eigen::Vector3d(p[4], p[5], p[6])));
    }
    for (size_t j = 0; j < points_.size(); ++j) {
      if (points_[j]->fixed()) continue;
      points_[j]->setEstimate(Eigen::Vector3d(xyz[3 * j], xyz[3 * j + 1], xyz[3 * j + 2]));
    }
    for (size_t e = 0; e < edges_.size(); ++e) {
      edges_[e]->error()[0] = err[2 * e];
      edges_[e]->error()[1] = err[2 * e + 1];
    }
    return true;
  }

  // set false and call writeBack() once after optimize() to skip the per-iteration read-back
  void setWriteBackEveryIteration(bool b) { write_back_every_iteration_ = b; }

  bool computeMarginals(g2o::SparseBlockMatrix<g2o::MatrixX> &, const std::vector<std::pair<int, int>> &) override { return false; }
  bool updateStructure(const std::vector<g2o::HyperGraph::Vertex *> &, const g2o::HyperGraph::EdgeSet &) override { return false; }

  // same fields OptimizationAlgorithmLevenberg::printVerbose writes (levenberg.cpp:187-193)
  void printVerbose(std::ostream &os) const override {
    os << "\t schur= 1\t lambda= " << last_.lambda << "\t levenbergIter= " << last_.trials;
  }
  double currentLambda() const { return last_.lambda; }
  int levenbergIteration() const { return last_.trials; }
  ssba_handle *handle() { return h_; }

 private:
  ssba_options opt_{};
  ssba_handle *h_ = nullptr;
  ssba_iter_record last_{};
  bool write_back_every_iteration_ = true;
  std::vector<VertexPoseT *> poses_;
  std::vector<VertexXYZT *> points_;
  std::vector<EdgeProjectionT *> edges_;
};

}  // namespace ssba

// Convenience alias for ssvio's own types when ssvio/g2otypes.hpp has been included first.
#ifdef SSVIO_G2OTYPES_HPP
namespace ssba {
template <class EdgeAccess = AccessorEdgeAccess>
using OptimizationAlgorithmLevenbergCuda =
    OptimizationAlgorithmLevenbergCudaT<ssvio::VertexPose, ssvio::VertexXYZ, ssvio::EdgeProjection, EdgeAccess>;
}
#endif
