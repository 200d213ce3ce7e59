"""Round-2 additions of the C ABI, on the GPU: structure reuse for an unchanged topology (rounds 2..5 of
src/ssvio/backend.cpp:175-203), the per-edge outlier mask of the culling step (backend.cpp:205-227), the force-stop
flag (thirdparty/g2o/g2o/core/sparse_optimizer.h:183-187, optimization_algorithm_levenberg.cpp:145), the remaining
G2OBatchStatistics fields (thirdparty/g2o/g2o/core/batch_stats.h:40-77), and buffer reuse across the widened rows."""
import threading
import time

import numpy as np
import pytest

from common import golden_case, rel
from ssvio_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def BA(ssba_lib):
    from ssvio_b200 import ba
    return ba.BundleAdjuster


def perturbed(g, seed):
    """Same topology, other values: estimates, measurements."""
    rng = np.random.default_rng(seed)
    return synth.Graph(K=g.K, ext=g.ext, poses=g.poses.copy(), pose_fixed=g.pose_fixed, points=g.points + 0.01 * rng.normal(size=g.points.shape),
                       point_fixed=g.point_fixed, pose_idx=g.pose_idx, point_idx=g.point_idx, cam_idx=g.cam_idx,
                       uv=g.uv + 0.3 * rng.normal(size=g.uv.shape), huber_delta=g.huber_delta, iters=g.iters)


@pytest.mark.parametrize("name", ["small_fixed", "cfg1"])
def test_structure_reuse_gives_the_same_result_as_a_fresh_handle(BA, name):
    g, _ = golden_case(name)
    g2 = perturbed(g, 7)
    with BA() as fresh:
        fresh.set_graph(g2)
        want = fresh.optimize(g.iters); want_p, want_x, want_e = fresh.poses(), fresh.points(), fresh.edge_errors()
    with BA() as opt:
        opt.set_graph(g)
        opt.optimize(g.iters)
        assert (opt.problem_info().n_structure_builds, opt.problem_info().n_structure_reuses) == (1, 0)
        opt.set_graph(g2)                     # values only
        got = opt.optimize(g.iters)
        assert (opt.problem_info().n_structure_builds, opt.problem_info().n_structure_reuses) == (1, 1)
        assert [t[2] for t in got.trace()] == [t[2] for t in want.trace()]
        for (c, lam, _), (cw, lw, _) in zip(got.trace(), want.trace()):
            assert rel(c, cw) < 1e-11 and rel(lam, lw) < 1e-9
        assert rel(got.chi2_robust, want.chi2_robust) < 1e-11
        np.testing.assert_allclose(opt.poses(), want_p, rtol=0, atol=1e-9)
        np.testing.assert_allclose(opt.points(), want_x, rtol=0, atol=1e-8)
        np.testing.assert_allclose(opt.edge_errors(), want_e, rtol=0, atol=1e-7)
        # a topology change (one more fixed landmark) rebuilds; so does an explicit drop
        g3 = perturbed(g, 8); g3.point_fixed = g.point_fixed.copy(); g3.point_fixed[3] ^= 1
        opt.set_graph(g3); opt.optimize(2)
        assert opt.problem_info().n_structure_builds == 2
        opt.set_graph(g3); opt.drop_structure(); opt.optimize(2)
        assert opt.problem_info().n_structure_builds == 3
        # permuted edges = another topology
        perm = np.random.default_rng(0).permutation(g.n_edges)
        g4 = synth.Graph(K=g.K, ext=g.ext, poses=g.poses, pose_fixed=g.pose_fixed, points=g.points, point_fixed=g.point_fixed,
                         pose_idx=g.pose_idx[perm].copy(), point_idx=g.point_idx[perm].copy(), cam_idx=g.cam_idx[perm].copy(),
                         uv=g.uv[perm].copy(), huber_delta=g.huber_delta, iters=g.iters)
        opt.set_graph(g4); r4 = opt.optimize(g.iters)
        assert opt.problem_info().n_structure_builds == 4
        with BA() as f2:
            f2.set_graph(g); r0 = f2.optimize(g.iters)
        assert rel(r4.chi2_robust, r0.chi2_robust) < 1e-11


def test_outlier_mask_matches_edge_errors(BA):
    g, _ = golden_case("small_fixed")
    with BA() as opt:
        opt.set_graph(g)
        opt.optimize(g.iters)
        err = opt.edge_errors()
        mask, n_out = opt.outlier_mask(5.891)
        want = ((err ** 2).sum(1) > 5.891).astype(np.uint8)
        np.testing.assert_array_equal(mask, want)
        assert n_out == int(want.sum()) == opt.count_outliers(5.891)[0]
        assert 0 < n_out < g.n_edges


def test_force_stop_flag(BA):
    g = synth.make_config("cfg2")
    with BA() as opt:
        opt.set_graph(g)
        opt.initialize_optimization()
        # raised before the call: no iteration starts (sparse_optimizer.cpp:388), the estimate stays
        opt.request_stop()
        rep = opt.optimize(10)
        assert rep.iterations == 0
        np.testing.assert_array_equal(opt.poses(), g.poses)
        opt.clear_stop()
        full = opt.optimize(10)
        assert full.iterations == 10
        # raised from another thread while a long optimize runs: it ends early, with a consistent state
        opt.reset_state()
        t = threading.Timer(0.002, opt.request_stop)
        t.start()
        rep = opt.optimize(120)
        t.join()
        opt.clear_stop()
        assert 0 <= rep.iterations < 120, rep.iterations
        chi = opt.chi2()[1]
        assert rel(chi, rep.chi2_robust) < 1e-12 and np.isfinite(opt.poses()).all()


def test_batch_statistics_fields(BA):
    g, _ = golden_case("cfg1")
    with BA() as opt:
        opt.set_graph(g)
        opt.profile_reset()
        rep = opt.optimize(g.iters)
        p, info = opt.profile(), opt.problem_info()
        assert p.outer_iterations == rep.iterations
        assert p.levenberg_iterations == sum(t[2] for t in rep.trace())
        assert p.hessian_pose_dimension == 6 * info.n_free_poses and p.hessian_landmark_dimension == 3 * info.n_free_points
        assert p.cholesky_nnz == 36 * (info.n_factor_blocks - info.n_free_poses) + 21 * info.n_free_poses
        assert p.ms_structure_build > 0 and 0 < p.ms_symbolic_decomposition


def test_pose_graph_and_pose_only_buffers_are_independent(BA):
    """pose_graph, then a pose_only call that grows its buffer, then pose_graph again on one handle."""
    pg = synth.make_pose_graph(60, seed=3)
    with BA() as opt:
        p1, r1 = opt.pose_graph_optimize(pg, iters=5)
        for nf in (4, 64, 700):  # growing batches: the pose-only buffer is re-allocated
            opt.pose_only_optimize(synth.make_pose_only(nf, 80, seed=nf))
        p2, r2 = opt.pose_graph_optimize(pg, iters=5)
        assert r1.iterations == r2.iterations and rel(r2.chi2_robust, r1.chi2_robust) < 1e-9
        np.testing.assert_allclose(p1, p2, atol=1e-9)
