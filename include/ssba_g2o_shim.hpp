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

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <mutex>
#include <thread>
#include <typeinfo>
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
  // backend.cpp:83 makes a new solver for every window; the handle (device arena, pinned staging, stream,
  // resident structure) is parked here for the next solver instead of being torn down and re-made each time
  ~OptimizationAlgorithmLevenbergCudaT() override { park(h_, opt_); }

  // OptimizationAlgorithmWithHessian::init (optimization_algorithm_with_hessian.cpp:48-73).  g2o calls it from
  // every optimize() (sparse_optimizer.cpp:379), i.e. once per round of backend.cpp:175-203, so it is written to be
  // cheap: exact-type checks instead of dynamic_cast on the 2e5 edges, the row of a vertex kept in the vertex
  // (colInHessian, which only g2o's own BlockSolver uses), flat arrays reused between calls.  When the active
  // graph is the one of the previous call, libssba keeps its structure and only takes the values.
  bool init(bool /*online*/ = false) override {
    if (!_optimizer) return false;
    const auto t_init0 = std::chrono::steady_clock::now();
    if (!h_) h_ = unpark(opt_);
    if (!h_ && ssba_create(&opt_, &h_) != SSBA_OK) {
      std::cerr << "ssba: " << ssba_last_error(nullptr) << std::endl;
      return false;
    }
    const auto &av = _optimizer->activeVertices();  // sorted by id (sparse_optimizer.cpp:493-498)
    const auto &ae = _optimizer->activeEdges();     // internalId = addEdge order
    edges_.clear();
    {
      // The vertices are 2e4 heap objects: two passes over a few threads - (1) the kind of every vertex and the
      // number of poses / landmarks per chunk, (2) with the chunk offsets known, rows (colInHessian), pointers,
      // estimates and fixed flags.  Rows follow the order of activeVertices() (by id), like g2o's own index mapping.
      const size_t nv = av.size();
      const int ntv = thread_count(4 * nv);
      kind_.resize(nv);
      std::vector<size_t> n_pose(ntv + 1, 0), n_point(ntv + 1, 0);
      std::vector<int> vbad(ntv, 0);
      parallel_chunks(nv, ntv, [&](int t, size_t i0, size_t i1) {
        size_t np = 0, nl = 0;
        for (size_t i = i0; i < i1; ++i) {
          if (i + 8 < i1) __builtin_prefetch(av[i + 8]);
          auto *v = av[i];
          const std::type_info &ti = typeid(*v);  // exact types first: a failing dynamic_cast walks the hierarchy
          if (ti == typeid(VertexXYZT)) { kind_[i] = 1; ++nl; }
          else if (ti == typeid(VertexPoseT)) { kind_[i] = 0; ++np; }
          else if (dynamic_cast<VertexPoseT *>(v)) { kind_[i] = 0; ++np; }
          else if (dynamic_cast<VertexXYZT *>(v)) { kind_[i] = 1; ++nl; }
          else { vbad[t] = 1; return; }
        }
        n_pose[t + 1] = np; n_point[t + 1] = nl;
      });
      for (int t = 0; t < ntv; ++t) {
        if (vbad[t]) { std::cerr << "ssba: unsupported vertex type in the active graph" << std::endl; return false; }
        n_pose[t + 1] += n_pose[t]; n_point[t + 1] += n_point[t];
      }
      const size_t NPose = n_pose[ntv], NPoint = n_point[ntv];
      poses_.resize(NPose); points_.resize(NPoint);
      pose_qt_.resize(7 * NPose); xyz_.resize(3 * NPoint); pose_fixed_.resize(NPose); point_fixed_.resize(NPoint);
      parallel_chunks(nv, ntv, [&](int t, size_t i0, size_t i1) {
        size_t ip = n_pose[t], il = n_point[t];
        for (size_t i = i0; i < i1; ++i) {
          if (i + 8 < i1) { const char *nx = reinterpret_cast<const char *>(av[i + 8]); __builtin_prefetch(nx); __builtin_prefetch(nx + 64); __builtin_prefetch(nx + 128); }
          if (kind_[i] == 0) {
            VertexPoseT *vp = static_cast<VertexPoseT *>(av[i]);
            vp->setColInHessian((int)ip);
            poses_[ip] = vp;
            const auto &T = vp->estimate();
            const auto &q = T.unit_quaternion();
            double *qt = &pose_qt_[7 * ip];
            qt[0] = q.x(); qt[1] = q.y(); qt[2] = q.z(); qt[3] = q.w();
            qt[4] = T.translation()[0]; qt[5] = T.translation()[1]; qt[6] = T.translation()[2];
            pose_fixed_[ip] = vp->fixed();
            ++ip;
          } else {
            VertexXYZT *vl = static_cast<VertexXYZT *>(av[i]);
            vl->setColInHessian((int)il);
            points_[il] = vl;
            const auto &p = vl->estimate();
            xyz_[3 * il] = p[0]; xyz_[3 * il + 1] = p[1]; xyz_[3 * il + 2] = p[2];
            point_fixed_[il] = vl->fixed();
            ++il;
          }
        }
      });
    }
    const auto t_init_v = std::chrono::steady_clock::now();
    const size_t ne = ae.size();
    if (ne == 0) return false;
    pidx_.resize(ne); lidx_.resize(ne); cam_.resize(ne); uv_.resize(2 * ne); info_.resize(3 * ne); delta_.resize(ne);
    edges_.resize(ne);
    // The edges are 2e5 heap objects: walking them is the bulk of init(), so it is spread over a few threads.
    // Every thread numbers the distinct camera extrinsics it meets locally; the lists are merged afterwards.
    const int nt = thread_count(ne);
    std::vector<std::vector<double>> ext_local(nt);
    std::vector<int> bad(nt, 0);
    double K[9] = {0};
    if (auto *e0 = dynamic_cast<const EdgeProjectionT *>(ae[0])) {  // one K for the whole graph (g2otypes.hpp:118-121)
      const Eigen::Matrix3d &Ke = EdgeAccess::K(e0);
      for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) K[3 * r + c] = Ke(r, c);
    }
    parallel_chunks(ne, nt, [&](int t, size_t i0, size_t i1) {
      std::vector<double> &ext = ext_local[t];
      for (size_t i = i0; i < i1; ++i) {
        // the walk is bound by cache misses (an edge object spans ~10 lines, its vertex array and robust kernel are
        // separate allocations): fetch the object 16 edges ahead, and what it points to 8 edges ahead
        if (i + 16 < i1) prefetch_edge(ae[i + 16]);
        if (i + 8 < i1) {
          const g2o::OptimizableGraph::Edge *e8 = ae[i + 8];
          __builtin_prefetch(e8->vertices().data());
          if (const auto *rk = e8->robustKernel()) { __builtin_prefetch(rk); __builtin_prefetch(reinterpret_cast<const char *>(rk) + sizeof(g2o::RobustKernelHuber) - 8); }
        }
        auto *e = ae[i];
        EdgeProjectionT *ep = typeid(*e) == typeid(EdgeProjectionT) ? static_cast<EdgeProjectionT *>(e) : dynamic_cast<EdgeProjectionT *>(e);
        if (!ep) { bad[t] = 1; return; }
        edges_[i] = ep;
        pidx_[i] = static_cast<const g2o::OptimizableGraph::Vertex *>(ep->vertex(0))->colInHessian();
        lidx_[i] = static_cast<const g2o::OptimizableGraph::Vertex *>(ep->vertex(1))->colInHessian();
        const auto &X = EdgeAccess::ext(ep);
        const auto &q = X.unit_quaternion();
        const double qt[7] = {q.x(), q.y(), q.z(), q.w(), X.translation()[0], X.translation()[1], X.translation()[2]};
        int ci = -1;
        for (size_t c = 0; c < ext.size() / 7; ++c)
          if (std::memcmp(&ext[7 * c], qt, sizeof(qt)) == 0) { ci = (int)c; break; }
        if (ci < 0) {
          if (ext.size() / 7 >= SSBA_MAX_CAMERAS) { bad[t] = 2; return; }
          ci = (int)(ext.size() / 7);
          ext.insert(ext.end(), qt, qt + 7);
        }
        cam_[i] = (uint8_t)ci;
        uv_[2 * i] = ep->measurement()[0]; uv_[2 * i + 1] = ep->measurement()[1];
        const auto &I = ep->information();
        info_[3 * i] = I(0, 0); info_[3 * i + 1] = I(0, 1); info_[3 * i + 2] = I(1, 1);
        double d = 0.0;
        if (auto *rk = ep->robustKernel()) {
          auto *hub = typeid(*rk) == typeid(g2o::RobustKernelHuber) ? static_cast<g2o::RobustKernelHuber *>(rk) : dynamic_cast<g2o::RobustKernelHuber *>(rk);
          if (!hub) { bad[t] = 3; return; }
          d = hub->delta();
        }
        delta_[i] = d;
      }
    });
    for (int t = 0; t < nt; ++t) {
      if (bad[t] == 1) { std::cerr << "ssba: unsupported edge type in the active graph" << std::endl; return false; }
      if (bad[t] == 3) { std::cerr << "ssba: only RobustKernelHuber is supported" << std::endl; return false; }
    }
    // merge the per-thread extrinsics (in thread order, i.e. in edge order) and renumber
    ext_qt_.clear();
    std::vector<std::vector<uint8_t>> remap(nt);
    for (int t = 0; t < nt; ++t) {
      for (size_t c = 0; c < ext_local[t].size() / 7; ++c) {
        int ci = -1;
        for (size_t k = 0; k < ext_qt_.size() / 7; ++k)
          if (std::memcmp(&ext_qt_[7 * k], &ext_local[t][7 * c], 7 * sizeof(double)) == 0) { ci = (int)k; break; }
        if (ci < 0) {
          if (bad[t] == 2 || ext_qt_.size() / 7 >= SSBA_MAX_CAMERAS) { std::cerr << "ssba: too many distinct camera extrinsics" << std::endl; return false; }
          ci = (int)(ext_qt_.size() / 7);
          ext_qt_.insert(ext_qt_.end(), &ext_local[t][7 * c], &ext_local[t][7 * c] + 7);
        }
        remap[t].push_back((uint8_t)ci);
      }
      if (bad[t] == 2) { std::cerr << "ssba: too many distinct camera extrinsics" << std::endl; return false; }
    }
    parallel_chunks(ne, nt, [&](int t, size_t i0, size_t i1) {
      for (size_t i = i0; i < i1; ++i) cam_[i] = remap[t][cam_[i]];
    });
    const auto t_init1 = std::chrono::steady_clock::now();
    bool ok = ssba_set_cameras(h_, K, (int32_t)(ext_qt_.size() / 7), ext_qt_.data()) == SSBA_OK &&
              ssba_set_poses(h_, (int32_t)poses_.size(), pose_qt_.data(), pose_fixed_.data()) == SSBA_OK &&
              ssba_set_points(h_, (int32_t)points_.size(), xyz_.data(), point_fixed_.data()) == SSBA_OK &&
              ssba_set_edges(h_, (int32_t)ne, pidx_.data(), lidx_.data(), cam_.data(), uv_.data(), info_.data(), delta_.data(), 0.0) == SSBA_OK &&
              ssba_initialize(h_) == SSBA_OK;
    if (!ok) std::cerr << "ssba: " << ssba_last_error(h_) << std::endl;
    if (timing()) {
      const auto t_init2 = std::chrono::steady_clock::now();
      std::fprintf(stderr, "[ssba shim] init: flatten %.3f ms (vertices %.3f), set_* + initialize %.3f ms\n", std::chrono::duration<double, std::milli>(t_init1 - t_init0).count(),
                   std::chrono::duration<double, std::milli>(t_init_v - t_init0).count(), std::chrono::duration<double, std::milli>(t_init2 - t_init1).count());
    }
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
    const auto t_wb0 = std::chrono::steady_clock::now();
    std::vector<double> &qt = wb_qt_, &xyz = wb_xyz_, &err = wb_err_;
    qt.resize(7 * poses_.size()); xyz.resize(3 * points_.size()); err.resize(2 * edges_.size());
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
    parallel_chunks(edges_.size(), thread_count(edges_.size()), [&](int, size_t e0, size_t e1) {
      for (size_t e = e0; e < e1; ++e) {
        if (e + 12 < e1) __builtin_prefetch(&edges_[e + 12]->error(), 1);
        edges_[e]->error()[0] = err[2 * e];
        edges_[e]->error()[1] = err[2 * e + 1];
      }
    });
    if (timing()) std::fprintf(stderr, "[ssba shim] writeBack %.3f ms\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_wb0).count());
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
  static bool timing() { static const bool on = std::getenv("SSBA_TIMING") != nullptr; return on; }
  // a few short-lived threads for the passes over the edge objects (SSBA_SHIM_THREADS overrides; 1 = none)
  static int thread_count(size_t n) {
    static const int cfg = [] { const char *e = std::getenv("SSBA_SHIM_THREADS"); return e ? std::atoi(e) : 0; }();
    int nt = cfg > 0 ? cfg : (int)std::min<unsigned>(8u, std::max(1u, std::thread::hardware_concurrency() / 2));
    if (n < 20000) nt = 1;
    return nt;
  }
  // the lines of an edge object init() reads: vptr / vertex container / robust kernel pointer at the front,
  // measurement + information behind the OptimizableGraph::Edge part, K / extrinsics at the end (addresses only)
  static void prefetch_edge(const g2o::OptimizableGraph::Edge *e) {
    const char *p = reinterpret_cast<const char *>(e);
    __builtin_prefetch(p); __builtin_prefetch(p + 64); __builtin_prefetch(p + 128); __builtin_prefetch(p + 192);
    const EdgeProjectionT *ep = static_cast<const EdgeProjectionT *>(e);  // never dereferenced here
    const char *x = reinterpret_cast<const char *>(&EdgeAccess::ext(ep));
    __builtin_prefetch(x); __builtin_prefetch(x + 56);
  }
  template <class F> static void parallel_chunks(size_t n, int nt, F &&fn) {
    if (nt <= 1) { fn(0, (size_t)0, n); return; }
    std::vector<std::thread> th;
    th.reserve(nt - 1);
    for (int t = 1; t < nt; ++t) th.emplace_back([&, t] { fn(t, n * t / nt, n * (t + 1) / nt); });
    fn(0, (size_t)0, n / nt);
    for (auto &x : th) x.join();
  }
  // one parked handle per process (header-only: function-local statics are shared across translation units)
  struct Parked { std::mutex mu; ssba_handle *h = nullptr; ssba_options opt{}; ~Parked() { if (h) ssba_destroy(h); } };
  static Parked &parked() { static Parked p; return p; }
  static void park(ssba_handle *h, const ssba_options &opt) {
    if (!h) return;
    Parked &p = parked();
    std::lock_guard<std::mutex> lk(p.mu);
    if (p.h) ssba_destroy(p.h);
    p.h = h; p.opt = opt;
  }
  static ssba_handle *unpark(const ssba_options &opt) {
    Parked &p = parked();
    std::lock_guard<std::mutex> lk(p.mu);
    if (!p.h || std::memcmp(&p.opt, &opt, sizeof(opt)) != 0) return nullptr;
    ssba_handle *h = p.h;
    p.h = nullptr;
    return h;
  }
  ssba_options opt_{};
  ssba_handle *h_ = nullptr;
  ssba_iter_record last_{};
  bool write_back_every_iteration_ = true;
  std::vector<VertexPoseT *> poses_;
  std::vector<VertexXYZT *> points_;
  std::vector<EdgeProjectionT *> edges_;
  // flat arrays handed to libssba (kept between calls: no allocation per round)
  std::vector<double> pose_qt_, xyz_, uv_, info_, delta_, ext_qt_, wb_qt_, wb_xyz_, wb_err_;
  std::vector<uint8_t> pose_fixed_, point_fixed_, cam_, kind_;
  std::vector<int32_t> pidx_, lidx_;
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
