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


# ---- failed factorisations (tests/golden/make_golden_nonpd.py, tests/test_gpu_nonpd.py)
# name -> (synth config, fraction of edges with information neg * I, neg, user lambda init, iterations)
NONPD_CASES = {
    "lambda30_small": ("small", 0.0, 0.0, 1e-30, 3),
    "lambda30_cfg2": ("cfg2", 0.0, 0.0, 1e-30, 3),
    "indef_cfg1": ("cfg1", 0.01, -1.0, 1e-3, 4),
    "indef_cfg2": ("cfg2", 0.01, -1.0, 1e-3, 3),
    "indef_cfg3": ("cfg3", 0.002, -1.0, 1e-3, 2),
}


def nonpd_case_inputs(name):
    """(graph, per-edge information or None, user lambda init, iterations) of a NONPD_CASES recipe."""
    cfg, frac, neg, ul, iters = NONPD_CASES[name]
    g = synth.make_config(cfg, seed=42)
    g.iters = iters
    info = None
    if frac > 0:
        rng = np.random.default_rng(5)
        info = np.tile(np.array([1.0, 0.0, 1.0]), (g.n_edges, 1))
        info[rng.random(g.n_edges) < frac] = np.array([neg, 0.0, neg])
        info = np.ascontiguousarray(info)
    return g, info, ul, iters
