"""Failed factorisations of the reduced pose system: "pivot <= 0 => solve() returns false"
(thirdparty/g2o/g2o/solvers/csparse/csparse_extension.cpp:115) => the trial is rejected with tempChi = DBL_MAX and
lambda is inflated (thirdparty/g2o/g2o/core/optimization_algorithm_levenberg.cpp:119-143), ten rejected trials end
the optimisation with Terminate (:147).  The CUDA path (both reduced solvers, cluster sizes 1 / 4 / 8) against traces
of the compiled reference (tests/golden/nonpd.json, made by tests/golden/make_golden_nonpd.py)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from common import GOLDEN, NONPD_CASES, nonpd_case_inputs, rel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_case(name):
    from ssvio_b200 import ba
    g, info, ul, iters = nonpd_case_inputs(name)
    with ba.BundleAdjuster(user_lambda_init=ul) as opt:
        opt.set_cameras(g.K, g.ext)
        opt.set_poses(g.poses, g.pose_fixed)
        opt.set_points(g.points, g.point_fixed)
        opt.set_edges(g.pose_idx, g.point_idx, g.cam_idx, g.uv, info=info, huber_delta_all=g.huber_delta)
        rep = opt.optimize(iters)
        poses = opt.poses()
        pi = opt.problem_info()
    return dict(iterations=rep.iterations, cholesky_failures=rep.cholesky_failures, chi2_initial=rep.chi2_initial,
                chi2_robust=rep.chi2_robust, lambda_final=rep.lambda_, last_result=rep.last_result,
                trace=[list(t) for t in rep.trace()], solver_kind=pi.solver_kind, solve_cluster=pi.solve_cluster,
                moved=float(np.abs(poses - g.poses).max()))


def check(name, got):
    want = json.load(open(os.path.join(GOLDEN, "nonpd.json")))[name]
    assert got["cholesky_failures"] > 0, got
    assert rel(got["chi2_initial"], want["chi2_initial"]) < 1e-11
    assert got["iterations"] == want["iterations"], (got, want)
    assert [t[2] for t in got["trace"]] == [t[2] for t in want["trace"]], (got["trace"], want["trace"])  # trials per iteration
    assert got["cholesky_failures"] == want["cholesky_failures"], (got, want)
    for (c, lam, _), (cw, lw, _) in zip(got["trace"], want["trace"]):
        assert rel(lam, lw) < 1e-6, (got["trace"], want["trace"])
        assert rel(c, cw) < 1e-6, (got["trace"], want["trace"])
    assert rel(got["chi2_robust"], want["chi2_robust"]) < 1e-6
    if name.startswith("lambda30"):
        # ten failed trials: Terminate (levenberg.cpp:147), optimize() stops after that iteration
        # (sparse_optimizer.cpp:388,401), and the estimate never moved
        assert got["last_result"] == 2 and got["iterations"] == 1 and got["moved"] == 0.0, got


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(NONPD_CASES))
def test_failed_factorisations_match_reference(ssba_lib, name):
    got = run_case(name)
    check(name, got)
    cfg = NONPD_CASES[name][0]
    assert got["solver_kind"] == 1, "the subtree-per-CTA solver should serve this size"
    assert got["solve_cluster"] == {"small": 1, "cfg1": 1, "cfg2": 4, "cfg3": 8}[cfg], got


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["lambda30_cfg2", "indef_cfg2", "indef_cfg3"])
def test_failed_factorisations_level_solver(ssba_lib, name):
    """The same through k_reduced_solve (the fallback solver; SSBA_SOLVER is read when the library loads)."""
    code = ("import json, sys; sys.path.insert(0, %r); sys.path.insert(0, %r); import test_gpu_nonpd as t; "
            "print('RESULT ' + json.dumps(t.run_case(%r)))" % (ROOT, os.path.join(ROOT, "tests"), name))
    env = dict(os.environ, SSBA_SOLVER="level")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    got = json.loads([l for l in r.stdout.splitlines() if l.startswith("RESULT ")][-1][7:])
    assert got["solver_kind"] == 0
    check(name, got)
