"""Generates the golden fixtures under tests/golden/ from the REFERENCE ITSELF.

Run in the build container (needs /root/reference):

    make -C oracle ref && python tests/golden/make_golden.py

The reference has no golden vectors or known-answer tests for this path, so
the fixtures are outputs of oracle/_ref/libssba_ref.so — the reference's own g2o core, CSparse
solver and ssvio g2otypes.hpp compiled from /root/reference (oracle/Makefile), driven by
oracle/ref_harness.cpp which replays src/ssvio/backend.cpp:81-203 on the synthetic graphs of
ssvio_b200/synth.py.  Each fixture holds the result of the path AS SHIPPED (numeric central
difference Jacobians) and with the corrected analytic Jacobians.

  *.npz   small cases: the full input arrays + reference outputs (estimates, per-edge errors)
  scalars.json   every case incl. the BASELINE configs: chi2 values and the per-iteration trace
                 (inputs are regenerated from (config, seed) by ssvio_b200.synth)
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.bindings import RefOracle  # noqa: E402
from ssvio_b200 import synth  # noqa: E402

# name -> (synth config, overrides, iterations, store inputs?)
CASES = {
    "tiny": ("tiny", {}, 5, True),
    "small": ("small", {}, 8, True),
    "small_fixed": ("small", dict(fix_first_pose=True, n_fixed_points=20), 8, True),
    "tiny_long": ("tiny", {}, 60, True),
    "cfg1": ("cfg1", {}, 5, True),
    "cfg2": ("cfg2", {}, 10, False),
    "cfg3": ("cfg3", {}, 10, False),
    "cfg5": ("cfg5", {}, 10, False),
}


def case_graph(name):
    cfg, over, iters, _ = CASES[name]
    g = synth.make_config(cfg, seed=42, **over)
    g.iters = iters
    return g


def main(argv):
    names = argv[1:] or list(CASES)
    ref = RefOracle()
    path = os.path.join(HERE, "scalars.json")
    scalars = json.load(open(path)) if os.path.exists(path) else {}
    for name in names:
        g = case_graph(name)
        store = CASES[name][3]
        entry = dict(n_poses=g.n_poses, n_points=g.n_points, n_edges=g.n_edges, iters=g.iters)
        arrays = {}
        for jac in ("numeric", "analytic"):
            r = ref.optimize(g, jacobian=jac, trace=True, want_state=store)
            rep = r["report"]
            entry[jac] = dict(
                iterations=rep.iterations, chi2_initial=rep.chi2_initial,
                chi2_robust=rep.chi2_robust, chi2_plain=rep.chi2_plain, lambda_=rep.lambda_,
                trace=[list(t) for t in rep.trace()])
            if store:
                arrays[f"{jac}_poses"] = r["poses"]
                arrays[f"{jac}_points"] = r["points"]
                arrays[f"{jac}_errors"] = r["errors"]
            print(f"{name:12s} {jac:8s} its={rep.iterations:3d} chi2={rep.chi2_robust:.10f}")
        scalars[name] = entry
        if store:
            np.savez_compressed(
                os.path.join(HERE, f"{name}.npz"), K=g.K, ext=g.ext, poses=g.poses,
                pose_fixed=g.pose_fixed, points=g.points, point_fixed=g.point_fixed,
                pose_idx=g.pose_idx, point_idx=g.point_idx, cam_idx=g.cam_idx, uv=g.uv,
                huber_delta=np.float64(g.huber_delta), iters=np.int32(g.iters), **arrays)
    json.dump(scalars, open(path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main(sys.argv)
