"""g2o text graphs <-> the flat local-BA layout of include/ssba.h.

Reads and writes the stock g2o records a monocular / per-camera bundle adjustment consists of,
in the form the reference's vendored g2o serialises them:

    VERTEX_SE3:EXPMAP id tx ty tz qx qy qz qw      camera-to-world, i.e. the INVERSE of the estimate
                                                   (VertexSE3Expmap::read/write,
                                                   g2o/types/sba/types_six_dof_expmap.cpp:93-108;
                                                   vector order SE3Quat::fromVector, se3quat.h:133-148)
    VERTEX_XYZ id x y z                            VertexSBAPointXYZ (g2o/types/sba/types_sba.cpp:40,180-186)
    EDGE_SE3_PROJECT_XYZ:EXPMAP idp idc u v i00 i01 i11
                                                   vertex 0 = point, vertex 1 = pose; upper triangle of
                                                   the 2x2 information (types_six_dof_expmap.cpp:363-387)
    FIX id                                         OptimizableGraph::load/saveVertex
                                                   (g2o/core/optimizable_graph.cpp:428,936)

Intrinsics are members of the edge in g2o and are not serialised, so the caller passes K (and,
for a stereo rig, which camera the file's edges belong to).  Vertex ids are kept only through
their order: poses and landmarks are numbered by increasing id, like
SparseOptimizer::initializeOptimization sorts them (g2o/core/sparse_optimizer.cpp:493-498);
edges keep the file order (= addEdge order = internalId).

Host-side tooling only: nothing here computes; the optimisation itself is libssba.
"""
from __future__ import annotations

import numpy as np

from . import synth

POSE_TAG, POINT_TAG, EDGE_TAG = "VERTEX_SE3:EXPMAP", "VERTEX_XYZ", "EDGE_SE3_PROJECT_XYZ:EXPMAP"


def _invert(q_xyzw: np.ndarray, t: np.ndarray):
    """Inverse of the rigid transform (q, t): (q*, -R(q*) t)."""
    qi = np.array([-q_xyzw[0], -q_xyzw[1], -q_xyzw[2], q_xyzw[3]]) / np.linalg.norm(q_xyzw)
    v, w = qi[:3], qi[3]
    uv = np.cross(v, t)
    rt = t + 2.0 * (w * uv + np.cross(v, uv))
    return qi, -rt


def load_g2o(path: str, K, ext=None, cam: int = 0, huber_delta: float = synth.HUBER_DELTA,
             iters: int = 10):
    """Parse a g2o text file into (Graph, info) — `info` is an (E, 3) array of the edges'
    information upper triangles, or None when every edge has the identity (what ssvio uses,
    src/ssvio/backend.cpp:161).  Unknown record types raise ValueError (nothing is dropped silently)."""
    poses, points, fixed, edges = {}, {}, set(), []
    with open(path) as f:
        for ln, line in enumerate(f, 1):
            tok = line.split()
            if not tok or tok[0].startswith("#"):
                continue
            if tok[0] == POSE_TAG:
                v = np.array(tok[2:9], dtype=np.float64)
                if v.size != 7:
                    raise ValueError(f"{path}:{ln}: {POSE_TAG} needs 7 numbers")
                q_cw, t_cw = _invert(v[3:7], v[0:3])
                poses[int(tok[1])] = np.concatenate([q_cw, t_cw])
            elif tok[0] == POINT_TAG:
                points[int(tok[1])] = np.array(tok[2:5], dtype=np.float64)
            elif tok[0] == EDGE_TAG:
                edges.append((int(tok[1]), int(tok[2]), [float(x) for x in tok[3:8]]))
            elif tok[0] == "FIX":
                fixed.update(int(x) for x in tok[1:])
            else:
                raise ValueError(f"{path}:{ln}: unsupported record '{tok[0]}'")
    pose_ids, point_ids = sorted(poses), sorted(points)
    prow = {i: r for r, i in enumerate(pose_ids)}
    lrow = {i: r for r, i in enumerate(point_ids)}
    ne = len(edges)
    pose_idx, point_idx = np.empty(ne, np.int32), np.empty(ne, np.int32)
    uv, info = np.empty((ne, 2)), np.empty((ne, 3))
    for e, (idp, idc, num) in enumerate(edges):
        if idp not in lrow or idc not in prow:
            raise ValueError(f"{path}: edge {e} references an unknown vertex ({idp}, {idc})")
        point_idx[e], pose_idx[e] = lrow[idp], prow[idc]
        uv[e], info[e] = num[0:2], num[2:5]
    K = np.asarray(K, dtype=np.float64).reshape(9)
    if ext is None:
        ext = np.array([[0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0]])
    g = synth.Graph(K=K, ext=np.asarray(ext, dtype=np.float64).reshape(-1, 7),
                    poses=np.array([poses[i] for i in pose_ids]).reshape(-1, 7),
                    pose_fixed=np.array([i in fixed for i in pose_ids], dtype=np.uint8),
                    points=np.array([points[i] for i in point_ids]).reshape(-1, 3),
                    point_fixed=np.array([i in fixed for i in point_ids], dtype=np.uint8),
                    pose_idx=pose_idx, point_idx=point_idx, cam_idx=np.full(ne, cam, np.uint8), uv=uv,
                    huber_delta=huber_delta, name=path, iters=iters)
    identity = ne == 0 or bool(np.all(info == np.array([1.0, 0.0, 1.0])))
    return g, (None if identity else info)


def save_g2o(path: str, g, cam: int = 0, info=None):
    """Write the poses, landmarks and the edges of camera `cam` of a Graph as g2o text (pose ids
    0..NK-1, landmark ids NK.., like backend.cpp:96-134 numbers them).  17 significant digits:
    the file round-trips the doubles exactly."""
    nk = g.n_poses
    fmt = lambda a: " ".join(repr(float(x)) for x in a)
    with open(path, "w") as f:
        for i in range(nk):
            q_wc, t_wc = _invert(g.poses[i, 0:4], g.poses[i, 4:7])
            f.write(f"{POSE_TAG} {i} {fmt(t_wc)} {fmt(q_wc)}\n")
            if g.pose_fixed[i]:
                f.write(f"FIX {i}\n")
        for j in range(g.n_points):
            f.write(f"{POINT_TAG} {nk + j} {fmt(g.points[j])}\n")
            if g.point_fixed[j]:
                f.write(f"FIX {nk + j}\n")
        for e in np.nonzero(g.cam_idx == cam)[0]:
            w = info[e] if info is not None else (1.0, 0.0, 1.0)
            f.write(f"{EDGE_TAG} {nk + int(g.point_idx[e])} {int(g.pose_idx[e])} {fmt(g.uv[e])} {fmt(w)}\n")
