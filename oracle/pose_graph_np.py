"""numpy restatement of the reference's pose-graph optimisation — TEST INFRASTRUCTURE (oracle "port" for
SURVEY.md 8f row 4; the CUDA path is not built yet).  Only tests/ may import it.

Follows:
  LoopClosing::PoseGraphOptimization        src/ssvio/loopclosing.cpp:458-532 (vertices, fixed rule as an
                                            input flag, identity 6x6 information, no robust kernel, optimize(20))
  EdgePoseGraph::computeError               include/ssvio/g2otypes.hpp:169-176: e = log(M^-1 * v0 * v1^-1)
  numeric Jacobians AS SHIPPED              g2o/core/base_binary_edge.hpp:144-212 (central differences,
                                            delta = 1e-9, through VertexPose::oplusImpl); jacobian="analytic"
                                            uses exact derivatives of the same error instead (for a tight check)
  BaseBinaryEdge::constructQuadraticForm    g2o/core/base_binary_edge.hpp:61-134
  OptimizationAlgorithmLevenberg::solve     g2o/core/optimization_algorithm_levenberg.cpp:58-175
  SE3::log / SO3::logAndTheta               sophus/se3.hpp:223-256, sophus/so3.hpp:245-286
The linear solve is dense here (the reference: BlockSolver<6,6> + LinearSolverEigen = sparse LDLT): same
solution up to rounding.  Pinned against the compiled reference (oracle/_ref, ssba_ref_pose_graph) by
tests/golden/make_golden_pose_graph.py and tests/test_pose_graph_oracle.py.
"""
from __future__ import annotations

import numpy as np

from ssvio_b200 import synth

EPS = 1e-10


def se3_log(T):
    """(N, 7) poses -> (N, 6) tangents (upsilon, omega)."""
    T = np.atleast_2d(T)
    v, w = T[:, :3], T[:, 3]
    n2 = (v * v).sum(1)
    n = np.sqrt(n2)
    small = n2 < EPS * EPS
    n_s = np.where(small, 1.0, n)
    f = np.where(small, 2.0 / w - (2.0 / 3.0) * n2 / (w ** 3), 2.0 * np.arctan(n_s / w) / n_s)
    theta = np.where(small, 2.0 * n2 / w, f * n)
    om = f[:, None] * v
    t = T[:, 4:7]
    oxt = np.cross(om, t)
    ooxt = np.cross(om, oxt)
    th_s = np.where(np.abs(theta) < EPS, 1.0, theta)
    c = np.where(np.abs(theta) < EPS, 1.0 / 12.0,
                 (1.0 - th_s * np.cos(0.5 * th_s) / (2.0 * np.sin(0.5 * th_s))) / (th_s * th_s))
    ups = t - 0.5 * oxt + c[:, None] * ooxt
    return np.concatenate([ups, om], axis=1)


def errors(poses, v0, v1, meas):
    """e = log(M^-1 * T_v0 * T_v1^-1), (E, 6)."""
    return se3_log(synth.se3_mul(synth.se3_mul(synth.se3_inv(meas), poses[v0]), synth.se3_inv(poses[v1])))


def _oplus(T, d):
    return synth.se3_mul(synth.se3_exp(d), T)


def _numeric_jacobians(poses, v0, v1, meas):
    """(E, 6, 6) each, central differences through oplus, like base_binary_edge.hpp:144-212."""
    delta, ne = 1e-9, len(v0)
    J0, J1 = np.empty((ne, 6, 6)), np.empty((ne, 6, 6))
    minv = synth.se3_inv(meas)
    for which, J in ((0, J0), (1, J1)):
        for d in range(6):
            add = np.zeros((1, 6)); add[0, d] = delta
            res = []
            for sgn in (1.0, -1.0):
                a = poses[v0] if which else _oplus(poses[v0], np.repeat(sgn * add, ne, 0))
                b = _oplus(poses[v1], np.repeat(sgn * add, ne, 0)) if which else poses[v1]
                res.append(se3_log(synth.se3_mul(synth.se3_mul(minv, a), synth.se3_inv(b))))
            J[:, :, d] = (res[0] - res[1]) / (2 * delta)
    return J0, J1


def optimize(poses0, fixed, v0, v1, meas, iters=20, jacobian="numeric"):
    """Returns (poses, trace [(chi2, lambda, trials)], iterations, chi2_initial)."""
    poses = np.array(poses0, float)
    fixed = np.asarray(fixed, bool)
    v0, v1 = np.asarray(v0), np.asarray(v1)
    n = poses.shape[0]
    active_edge = ~(fixed[v0] & fixed[v1])                    # sparse_optimizer.cpp:237
    v0, v1, meas = v0[active_edge], v1[active_edge], np.asarray(meas, float)[active_edge]
    touched = np.zeros(n, bool); touched[v0] = True; touched[v1] = True
    free = np.nonzero(~fixed & touched)[0]
    col = -np.ones(n, int); col[free] = np.arange(len(free))
    nf = len(free)
    tau, good_lo, good_hi, max_trials = 1e-5, 1.0 / 3.0, 2.0 / 3.0, 10
    chi = lambda P: float((errors(P, v0, v1, meas) ** 2).sum())
    chi0 = chi(poses)
    trace, lam, ni, its = [], 0.0, 2.0, 0
    for it in range(iters):
        cur = chi(poses)
        e = errors(poses, v0, v1, meas)
        if jacobian == "numeric":
            J0, J1 = _numeric_jacobians(poses, v0, v1, meas)
        else:
            raise ValueError("only the reference's numeric Jacobians are restated")
        H = np.zeros((6 * nf, 6 * nf)); b = np.zeros(6 * nf)
        for k in range(len(v0)):
            for (va, Ja) in ((v0[k], J0[k]), (v1[k], J1[k])):
                ca = col[va]
                if ca < 0:
                    continue
                b[6 * ca:6 * ca + 6] -= Ja.T @ e[k]
                for (vb, Jb) in ((v0[k], J0[k]), (v1[k], J1[k])):
                    cb = col[vb]
                    if cb >= 0:
                        H[6 * ca:6 * ca + 6, 6 * cb:6 * cb + 6] += Ja.T @ Jb
        if it == 0:
            lam, ni = tau * np.abs(np.diag(H)).max(), 2.0
        rho, qmax, stop = 0.0, 0, False
        while True:
            try:
                L = np.linalg.cholesky(H + lam * np.eye(6 * nf))
                x = np.linalg.solve(L.T, np.linalg.solve(L, b)); ok = True
            except np.linalg.LinAlgError:
                x, ok = np.zeros(6 * nf), False
            trial = poses.copy()
            if ok:
                trial[free] = _oplus(poses[free], x.reshape(nf, 6))
            tmp = chi(trial) if ok else np.finfo(float).max
            rho = (cur - tmp) / (float((x * (lam * x + b)).sum()) + 1e-3)
            if rho > 0 and np.isfinite(tmp):
                lam *= max(good_lo, min(1.0 - (2 * rho - 1) ** 3, good_hi)); ni = 2.0
                poses, cur = trial, tmp
            else:
                lam *= ni; ni *= 2
                if not np.isfinite(lam):
                    stop = True
                    break
            qmax += 1
            if not (rho < 0 and qmax < max_trials):
                break
        its += 1
        trace.append((chi(poses), lam, qmax))
        if qmax == max_trials or rho == 0 or stop:
            break
    return poses, trace, its, chi0
