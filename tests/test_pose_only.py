"""Pose-only LM of the front-end (FrontEnd::EstimateCurrentPose, src/ssvio/frontend.cpp:184-260),
batched over frames.

Checkers: tests/golden/pose_only.npz (made by the compiled reference: the reference's own g2o +
EdgeProjectionPoseOnly, tests/golden/make_golden_pose_only.py) and the numpy restatement
(oracle/pose_only_np.py).  The reference uses its analytic Jacobian here, so the bar is tight:
same inlier / outlier decisions, chi2 to 1e-9 relative, poses to 1e-7.
"""
import os

import numpy as np
import pytest

from ssvio_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pose_only.npz")


def golden_batches():
    z = np.load(GOLD)
    for tag in ("a", "b", "c"):
        b = synth.PoseOnlyBatch(K=z[f"{tag}_K"], feat_ptr=z[f"{tag}_feat_ptr"], poses=z[f"{tag}_poses"],
                                xyz=z[f"{tag}_xyz"], uv=z[f"{tag}_uv"])
        yield tag, b, dict(poses=z[f"{tag}_ref_poses"], outlier=z[f"{tag}_ref_outlier"],
                           inliers=z[f"{tag}_ref_inliers"], chi2=z[f"{tag}_ref_chi2"],
                           # the loop closer's schedule (LoopClosing::OptimizeCurrentPose, loopclosing.cpp:245-351)
                           loop=dict(poses=z[f"{tag}_loop_poses"], outlier=z[f"{tag}_loop_outlier"],
                                     inliers=z[f"{tag}_loop_inliers"], chi2=z[f"{tag}_loop_chi2"]))


def check(tag, got, want, pose_atol=1e-7):
    poses, flags, n_in, chi = got
    np.testing.assert_array_equal(n_in, want["inliers"], err_msg=tag)
    np.testing.assert_array_equal(flags, want["outlier"], err_msg=tag)
    np.testing.assert_allclose(chi, want["chi2"], rtol=1e-9, err_msg=tag)
    np.testing.assert_allclose(poses, want["poses"], rtol=0, atol=pose_atol, err_msg=tag)


def test_numpy_restatement_matches_reference_fixtures():
    from oracle import pose_only_np
    for tag, b, want in golden_batches():
        check(tag, pose_only_np.optimize(b.K, b.feat_ptr, b.poses, b.xyz, b.uv), want)
        check(tag + " (loop closing)", pose_only_np.optimize(b.K, b.feat_ptr, b.poses, b.xyz, b.uv, pre_rounds=1), want["loop"])


def test_generator_is_reproducible():
    z = np.load(GOLD)
    b = synth.make_pose_only(16, 150, seed=3)
    np.testing.assert_array_equal(b.feat_ptr, z["a_feat_ptr"])
    np.testing.assert_array_equal(b.uv, z["a_uv"])
    np.testing.assert_array_equal(b.poses, z["a_poses"])


def test_outliers_are_found_and_poses_improve():
    """Properties that hold whatever the implementation: the gross outliers end up flagged and the
    pose gets closer to the truth the measurements were generated from."""
    from oracle import pose_only_np
    nf = 5
    b = synth.make_pose_only(nf, 120, seed=77, outlier_frac=0.15)
    poses, flags, n_in, _ = pose_only_np.optimize(b.K, b.feat_ptr, b.poses, b.xyz, b.uv)
    idx = np.arange(nf, dtype=np.float64)
    gt = synth.se3_exp(np.stack([0.02 * idx, 0.01 * idx, -idx, 0.002 * idx, 0.002 * idx, 0.002 * idx], axis=1))
    assert np.abs(poses[:, 4:] - gt[:, 4:]).max() < 0.25 * np.abs(b.poses[:, 4:] - gt[:, 4:]).max()
    assert 0.05 < flags.mean() < 0.3 and (n_in == np.diff(b.feat_ptr) - np.add.reduceat(flags, b.feat_ptr[:-1])).all()


@pytest.mark.gpu
def test_cuda_matches_reference_fixtures(ssba_lib):
    from ssvio_b200 import ba
    with ba.BundleAdjuster() as opt:
        for tag, b, want in golden_batches():
            check(tag, opt.pose_only_optimize(b), want)
            check(tag + " (loop closing)", opt.pose_only_optimize(b, loop_closing=True), want["loop"])


@pytest.mark.gpu
def test_cuda_matches_oracle_on_fresh_batches_and_edge_cases(ssba_lib):
    from oracle import pose_only_np
    from ssvio_b200 import ba
    with ba.BundleAdjuster() as opt:
        for seed, nf, nfeat, kw in [(101, 37, 90, {}), (102, 3, 500, dict(outlier_frac=0.25)), (103, 64, 20, {})]:
            b = synth.make_pose_only(nf, nfeat, seed=seed, **kw)
            want = dict(zip(("poses", "outlier", "inliers", "chi2"), pose_only_np.optimize(b.K, b.feat_ptr, b.poses, b.xyz, b.uv)))
            check(f"seed{seed}", opt.pose_only_optimize(b), want)
        # a frame without features keeps its pose; zero frames is a no-op
        b = synth.make_pose_only(3, 50, seed=5)
        fp = b.feat_ptr.copy(); fp[1:] = fp[1]  # frames 1 and 2 lose their features
        b2 = synth.PoseOnlyBatch(K=b.K, feat_ptr=fp, poses=b.poses, xyz=b.xyz[:fp[-1]], uv=b.uv[:fp[-1]])
        poses, flags, n_in, _ = opt.pose_only_optimize(b2)
        np.testing.assert_array_equal(poses[1:], b.poses[1:])
        assert n_in[1] == 0 and n_in[2] == 0 and n_in[0] > 0
        empty = synth.PoseOnlyBatch(K=b.K, feat_ptr=np.zeros(1, np.int32), poses=np.zeros((0, 7)), xyz=np.zeros((0, 3)), uv=np.zeros((0, 2)))
        assert opt.pose_only_optimize(empty)[0].shape == (0, 7)
        # the bundle-adjustment problem of the same handle is untouched
        g = synth.make_config("tiny")
        opt.set_graph(g)
        r1 = opt.optimize(g.iters)
        opt.pose_only_optimize(b)
        opt.reset_state()
        r2 = opt.optimize(g.iters)
        assert abs(r1.chi2_robust - r2.chi2_robust) <= 1e-12 * r1.chi2_robust
