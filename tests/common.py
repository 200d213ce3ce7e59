"""Shared helpers of the test-suite: golden fixtures -> Graph, tolerances."""
from __future__ import annotations

import json
import os

import numpy as np

from ssvio_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# north_star: final chi2 within 1e-6 relative of the reference g2o path
CHI2_RTOL = 1e-6
# the restatement / the CUDA path with the same (analytic) Jacobians agree far tighter
CHI2_RTOL_SAME_JACOBIAN = 1e-9


def golden_scalars():
    return json.load(open(os.path.join(GOLDEN, "scalars.json")))


def golden_case(name):
    """(Graph, npz) of a stored small case."""
    z = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    g = synth.Graph(K=z["K"], ext=z["ext"], poses=z["poses"], pose_fixed=z["pose_fixed"],
                    points=z["points"], point_fixed=z["point_fixed"], pose_idx=z["pose_idx"],
                    point_idx=z["point_idx"], cam_idx=z["cam_idx"], uv=z["uv"],
                    huber_delta=float(z["huber_delta"]), name=name, iters=int(z["iters"]))
    return g, z


def rel(a, b):
    return abs(a - b) / max(abs(b), 1e-300)


def converging_prefix(trace, tol=1e-9):
    """Iterations while chi2 still drops by more than `tol` relative: past that point the
    accept/reject decisions are rounding noise (SURVEY.md 7 'trajectory divergence')."""
    n = 0
    prev = None
    for chi, lam, trials in trace:
        if prev is not None and abs(prev - chi) <= tol * abs(prev):
            break
        prev = chi
        n += 1
    return n
