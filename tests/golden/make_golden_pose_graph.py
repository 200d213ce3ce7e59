"""Generates tests/golden/pose_graph.npz: seeded loop-closure pose graphs (ssvio_b200.synth.make_pose_graph)
and what the compiled reference (oracle/_ref, ssba_ref_pose_graph = LoopClosing::PoseGraphOptimization,
src/ssvio/loopclosing.cpp:458-532, with the reference's own g2o + EdgePoseGraph + LinearSolverEigen) returns.
Run in the build container (needs /root/reference):  python tests/golden/make_golden_pose_graph.py"""
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
for tag, (n, seed, kw) in {"a": (60, 5, {}), "b": (25, 11, dict(n_loops=1)), "c": (200, 23, dict(n_loops=6, drift_t=0.05))}.items():
    pg = synth.make_pose_graph(n, seed=seed, **kw)
    poses, rep = bindings.ref_pose_graph(ref.lib, pg)
    out.update({f"{tag}_poses": pg.poses, f"{tag}_fixed": pg.fixed, f"{tag}_v0": pg.v0, f"{tag}_v1": pg.v1, f"{tag}_meas": pg.meas,
                f"{tag}_ref_poses": poses, f"{tag}_ref_trace": np.array(rep.trace()), f"{tag}_ref_chi2": rep.chi2_robust,
                f"{tag}_ref_chi2_initial": rep.chi2_initial, f"{tag}_ref_iterations": rep.iterations})
    # the same optimisation stopped after 1 and 3 iterations: the numeric Jacobians' noise has not been amplified by
    # the soft modes of the chain yet, so these pin the poses tightly
    for it in (1, 3):
        p_it, r_it = bindings.ref_pose_graph(ref.lib, pg, iters=it)
        out.update({f"{tag}_ref_poses_it{it}": p_it, f"{tag}_ref_chi2_it{it}": r_it.chi2_robust})
    print(tag, "key-frames", n, "edges", len(pg.v0), "chi2", rep.chi2_initial, "->", rep.chi2_robust, "iterations", rep.iterations)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "pose_graph.npz"), **out)
