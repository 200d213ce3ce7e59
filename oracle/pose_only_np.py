"""numpy restatement of the reference's pose-only LM — TEST INFRASTRUCTURE (oracle "port" for
SURVEY.md 8f row 3).  Only tests/ may import it; the product is ssba_pose_only_optimize in libssba.

Follows, line by line:
  FrontEnd::EstimateCurrentPose             src/ssvio/frontend.cpp:184-260 (rounds, outlier rule,
                                            kernel removal at round rounds - 2)
  EdgeProjectionPoseOnly                    include/ssvio/g2otypes.hpp:67-110 (error :78-85, analytic
                                            Jacobian as shipped :87-101, Zinv = 1 / (Z + 1e-18))
  BaseUnaryEdge::constructQuadraticForm     g2o/core/base_unary_edge.hpp:49-79
  RobustKernelHuber (delta = 1)             g2o/core/robust_kernel_impl.cpp:65-78
  OptimizationAlgorithmLevenberg::solve     g2o/core/optimization_algorithm_levenberg.cpp:58-175
  SparseOptimizer::optimize                 g2o/core/sparse_optimizer.cpp:366-431
  LinearSolverDense (LDLT, isPositive)      g2o/solvers/dense/linear_solver_dense.h:56-118
  VertexPose::oplusImpl / SE3::exp          include/ssvio/g2otypes.hpp:36-41, sophus/se3.hpp:763-784
One frame at a time, vectorised over its features.  Pinned against the compiled reference
(oracle/_ref, ssba_ref_pose_only) by tests/golden/make_golden_pose_only.py and tests/test_pose_only.py.
"""
from __future__ import annotations

import numpy as np

from ssvio_b200 import synth


def _rotate(q, p):
    """Sophus SO3 point action (so3.hpp:352-360) for one quaternion (x, y, z, w), many points."""
    v, w = q[:3], q[3]
    uv = 2.0 * np.cross(v, p)
    return p + w * uv + np.cross(v, uv)


def _errors(K, T, xyz, uv):
    P = _rotate(T[:4], xyz) + T[4:7]
    pix = P @ K.reshape(3, 3).T
    return uv - pix[:, :2] / pix[:, 2:3], P


def _jacobians(K, P):
    fx, fy = K[0], K[4]
    X, Y, Z = P[:, 0], P[:, 1], P[:, 2]
    zi = 1.0 / (Z + 1e-18)
    zi2 = zi * zi
    J = np.zeros((P.shape[0], 2, 6))
    J[:, 0, 0] = -fx * zi; J[:, 0, 2] = fx * X * zi2; J[:, 0, 3] = fx * X * Y * zi2
    J[:, 0, 4] = -fx - fx * X * X * zi2; J[:, 0, 5] = fx * Y * zi
    J[:, 1, 1] = -fy * zi; J[:, 1, 2] = fy * Y * zi2; J[:, 1, 3] = fy + fy * Y * Y * zi2
    J[:, 1, 4] = -fy * X * Y * zi2; J[:, 1, 5] = -fy * X * zi
    return J


def _huber(e2, delta=1.0):
    rho0, rho1 = e2.copy(), np.ones_like(e2)
    m = e2 > delta * delta
    s = np.sqrt(e2[m])
    rho0[m] = 2 * s * delta - delta * delta
    rho1[m] = delta / s
    return rho0, rho1


def _robust_chi2(e, robust):
    e2 = (e * e).sum(1)
    return float((_huber(e2)[0] if robust else e2).sum())


def _oplus(T, d):
    """VertexPose::oplusImpl: T <- SE3::exp(d) * T (quaternion re-normalised by the product)."""
    return synth.se3_mul(synth.se3_exp(d)[0], T)


def optimize_frame(K, T0, xyz, uv, rounds=4, iters=10, chi2_threshold=5.991, pre_rounds=0):
    """Returns (T, outlier flags, n_inliers, robust chi2 after the last optimize).
    pre_rounds: initializeOptimization(); optimize(iters) calls BEFORE the classified rounds, with every edge
    and the kernel on: 0 = FrontEnd::EstimateCurrentPose, 1 = LoopClosing::OptimizeCurrentPose
    (src/ssvio/loopclosing.cpp:302-303)."""
    K = np.asarray(K, float).reshape(9)
    T = np.array(T0, float)
    n = xyz.shape[0]
    outlier = np.zeros(n, bool)
    level1 = np.zeros(n, bool)
    robust = True
    err = np.zeros((n, 2))          # the edges' _error members
    chi_last, cnt_out = 0.0, 0
    tau, good_lo, good_hi, max_trials = 1e-5, 1.0 / 3.0, 2.0 / 3.0, 10
    for rnd in range(-pre_rounds, rounds):
        act = ~level1                                   # initializeOptimization(): level-0 edges
        if act.any():
            lam, ni = 0.0, 2.0
            for it in range(iters):
                err[act], P = _errors(K, T, xyz[act], uv[act])      # computeActiveErrors
                cur = _robust_chi2(err[act], robust)
                J = _jacobians(K, P)
                e2 = (err[act] ** 2).sum(1)
                w = _huber(e2)[1] if robust else np.ones_like(e2)
                H = np.einsum("n,nki,nkj->ij", w, J, J)
                b = -np.einsum("n,nki,nk->i", w, J, err[act])
                if it == 0:
                    lam, ni = tau * np.abs(np.diag(H)).max(), 2.0
                rho, qmax, stop = 0.0, 0, False
                while True:
                    A = H + lam * np.eye(6)
                    try:
                        L = np.linalg.cholesky(A)
                        x = np.linalg.solve(L.T, np.linalg.solve(L, b))
                        ok = True
                    except np.linalg.LinAlgError:
                        x, ok = np.zeros(6), False
                    Tn = _oplus(T, x) if ok else T
                    err[act], _ = _errors(K, Tn, xyz[act], uv[act])
                    tmp = _robust_chi2(err[act], robust) if ok else np.finfo(float).max
                    scale = float((x * (lam * x + b)).sum()) + 1e-3
                    rho = (cur - tmp) / scale
                    if rho > 0 and np.isfinite(tmp):
                        alpha = min(1.0 - (2 * rho - 1) ** 3, good_hi)
                        lam *= max(good_lo, alpha)
                        ni = 2.0
                        cur = tmp
                        T = Tn
                    else:
                        lam *= ni
                        ni *= 2
                        if not np.isfinite(lam):
                            stop = True
                            break
                    qmax += 1
                    if not (rho < 0 and qmax < max_trials):
                        break
                if qmax == max_trials or rho == 0 or stop:
                    break                                # Terminate: optimize() leaves its loop
            chi_last = _robust_chi2(err[act], robust)    # of the _error members, like activeRobustChi2()
        if rnd < 0:
            continue                                     # loopclosing.cpp:302-303: no classification yet
        # frontend.cpp:243-262
        if outlier.any():
            err[outlier], _ = _errors(K, T, xyz[outlier], uv[outlier])
        chi = (err * err).sum(1)
        outlier = chi > chi2_threshold
        level1 = outlier.copy()
        cnt_out = int(outlier.sum())
        if rnd == rounds - 2:
            robust = False
    return T, outlier.astype(np.uint8), n - cnt_out, chi_last


def optimize(K, feat_ptr, poses, xyz, uv, rounds=4, iters=10, chi2_threshold=5.991, pre_rounds=0):
    nf = len(feat_ptr) - 1
    out_T, out_flag = np.empty((nf, 7)), np.zeros(xyz.shape[0], np.uint8)
    n_in, chi = np.zeros(nf, np.int32), np.zeros(nf)
    for f in range(nf):
        a, b = feat_ptr[f], feat_ptr[f + 1]
        out_T[f], out_flag[a:b], n_in[f], chi[f] = optimize_frame(K, poses[f], xyz[a:b], uv[a:b], rounds, iters, chi2_threshold, pre_rounds)
    return out_T, out_flag, n_in, chi
