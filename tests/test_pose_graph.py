"""Pose-graph optimisation on the B200 (ssba_pose_graph_optimize; LoopClosing::PoseGraphOptimization,
src/ssvio/loopclosing.cpp:458-532) against the fixtures of the compiled reference and the
numpy restatement.

What can be pinned how tightly: the reference differentiates numerically (delta = 1e-9, ~1e-7 relative noise on
every Jacobian entry, not reproducible bit for bit even between the reference and a restatement on the same CPU).
The noise moves the estimate along the SOFT modes of a long chain (directions the edges barely constrain) and
hardly at all along the constrained ones.  So: (1) the poses are compared after 1 and 3 iterations, before the
soft modes have amplified anything; (2) after the full optimize(20) the comparison is made where the graph does
constrain the estimate: the 6-vector error of every edge at the returned poses against the same at the reference's
poses (measured between reference and numpy restatement: 5e-7 on the 25 / 60 key-frame graphs, 5e-5 on the 200
key-frame one, where the poses themselves differ by 3e-2); (3) chi2: initial to 1e-11, trace prefix, final to 1e-6."""
import numpy as np
import pytest

from ssvio_b200 import synth
from test_pose_graph_oracle import GOLD, _compare_prefix

pytestmark = pytest.mark.gpu


def _graph(z, tag):
    return synth.PoseGraph(poses=z[f"{tag}_poses"], fixed=z[f"{tag}_fixed"], v0=z[f"{tag}_v0"], v1=z[f"{tag}_v1"], meas=z[f"{tag}_meas"])


# per graph: poses after 1 / 3 iterations (relative to the largest displacement), per-edge errors after optimize(20)
TOL = {"a": (2e-6, 5e-5, 2e-6), "b": (2e-6, 2e-5, 2e-6), "c": (5e-5, 5e-4, 2e-4)}


def test_cuda_matches_reference_fixtures(ssba_lib):
    from oracle import pose_graph_np
    from ssvio_b200 import ba
    z = np.load(GOLD)
    with ba.BundleAdjuster() as opt:
        for tag in ("a", "b", "c"):
            pg = _graph(z, tag)
            tol1, tol3, tol_e = TOL[tag]
            fx = pg.fixed.astype(bool)
            # (1) early iterations: poses
            for it, tol in ((1, tol1), (3, tol3)):
                poses, rep = opt.pose_graph_optimize(pg, iters=it)
                want = z[f"{tag}_ref_poses_it{it}"]
                moved = float(np.abs(want - pg.poses).max())
                np.testing.assert_allclose(poses, want, rtol=0, atol=tol * max(1.0, moved), err_msg=f"{tag} after {it} iterations")
                assert abs(rep.chi2_robust - float(z[f"{tag}_ref_chi2_it{it}"])) <= 1e-5 * float(z[f"{tag}_ref_chi2_it{it}"]), (tag, it)
            # (2) + (3) the full run
            poses, rep = opt.pose_graph_optimize(pg)
            assert abs(rep.chi2_initial - float(z[f"{tag}_ref_chi2_initial"])) <= 1e-11 * rep.chi2_initial, tag
            # the 200-key-frame graph amplifies the noise of the numeric Jacobians mid-trajectory (4e-5 seen)
            assert _compare_prefix(rep.trace(), z[f"{tag}_ref_trace"], tol=1e-5 if tag != "c" else 2e-4) >= 3, tag
            assert abs(rep.chi2_robust - float(z[f"{tag}_ref_chi2"])) <= 1e-6 * float(z[f"{tag}_ref_chi2"]), tag
            e_got = pose_graph_np.errors(poses, pg.v0, pg.v1, pg.meas)
            e_ref = pose_graph_np.errors(z[f"{tag}_ref_poses"], pg.v0, pg.v1, pg.meas)
            np.testing.assert_allclose(e_got, e_ref, rtol=0, atol=tol_e, err_msg=f"{tag}: per-edge errors at the final poses")
            # the soft modes themselves: bounded, not pinned (see the module docstring)
            moved = float(np.abs(z[f"{tag}_ref_poses"] - pg.poses).max())
            np.testing.assert_allclose(poses, z[f"{tag}_ref_poses"], rtol=0, atol=(1e-4 if tag != "c" else 1e-2) * max(1.0, moved), err_msg=tag)
            np.testing.assert_array_equal(poses[fx], pg.poses[fx])


def test_cuda_matches_oracle_and_edge_cases(ssba_lib):
    from oracle import pose_graph_np
    from ssvio_b200 import ba
    with ba.BundleAdjuster() as opt:
        pg = synth.make_pose_graph(40, seed=77, n_loops=2)
        want_poses, want_trace, _, chi0 = pose_graph_np.optimize(pg.poses, pg.fixed, pg.v0, pg.v1, pg.meas)
        poses, rep = opt.pose_graph_optimize(pg)
        assert abs(rep.chi2_initial - chi0) <= 1e-11 * chi0
        assert _compare_prefix(rep.trace(), want_trace) >= 3
        assert abs(rep.chi2_robust - want_trace[-1][0]) <= 1e-6 * want_trace[-1][0]
        np.testing.assert_allclose(poses, want_poses, rtol=0, atol=1e-3)
        # everything fixed: nothing to optimise, like optimize() returning -1
        allfixed = synth.PoseGraph(poses=pg.poses, fixed=np.ones_like(pg.fixed), v0=pg.v0, v1=pg.v1, meas=pg.meas)
        poses, rep = opt.pose_graph_optimize(allfixed)
        assert rep.iterations == -1
        np.testing.assert_array_equal(poses, pg.poses)
        # a long chain (nested-dissection order, cluster solver)
        big = synth.make_pose_graph(400, seed=9, n_loops=8)
        poses, rep = opt.pose_graph_optimize(big, iters=10)
        assert rep.iterations == 10 and rep.chi2_robust < 0.01 * rep.chi2_initial and rep.cholesky_failures == 0
        # H and b are assembled in a fixed order (k_pg_assemble, no atomics): the same bits again
        poses2, rep2 = opt.pose_graph_optimize(big, iters=10)
        np.testing.assert_array_equal(poses, poses2)
        assert rep2.trace() == rep.trace()
