"""Generates tests/golden/pose_only.npz: seeded pose-only batches (ssvio_b200.synth.make_pose_only)
and what the compiled reference (oracle/_ref, ssba_ref_pose_only = FrontEnd::EstimateCurrentPose,
src/ssvio/frontend.cpp:184-260, with the reference's own g2o + EdgeProjectionPoseOnly) returns
for them.  Run in the build container (needs /root/reference):  python tests/golden/make_golden_pose_only.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bindings  # noqa: E402
from ssvio_b200 import synth  # noqa: E402

bindings.build(which=("ref",))
ref = bindings.RefOracle()
out = {}
for tag, (nf, nfeat, seed, kw) in {"a": (16, 150, 3, {}), "b": (6, 40, 9, dict(outlier_frac=0.3)),
                                   "c": (4, 300, 21, dict(pose_sigma_t=0.3, pose_sigma_r=0.03))}.items():
    b = synth.make_pose_only(nf, nfeat, seed=seed, **kw)
    poses, flags, n_in, chi = bindings.ref_pose_only(ref.lib, b)
    # the loop closer's schedule (LoopClosing::OptimizeCurrentPose, src/ssvio/loopclosing.cpp:245-351)
    lposes, lflags, ln_in, lchi = bindings.ref_pose_only(ref.lib, b, pre_rounds=1)
    out.update({f"{tag}_loop_poses": lposes, f"{tag}_loop_outlier": lflags, f"{tag}_loop_inliers": ln_in, f"{tag}_loop_chi2": lchi})
    out.update({f"{tag}_K": b.K, f"{tag}_feat_ptr": b.feat_ptr, f"{tag}_poses": b.poses, f"{tag}_xyz": b.xyz,
                f"{tag}_uv": b.uv, f"{tag}_ref_poses": poses, f"{tag}_ref_outlier": flags,
                f"{tag}_ref_inliers": n_in, f"{tag}_ref_chi2": chi})
    print(tag, "frames", nf, "features", b.xyz.shape[0], "inliers", n_in.tolist())
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "pose_only.npz"), **out)
