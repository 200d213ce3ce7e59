"""ctypes bindings of the parity oracle — TEST INFRASTRUCTURE.

Two checkers live here, neither is part of the product:
  * RefOracle  — oracle/_ref/libssba_ref.so: the reference's own g2o/CSparse + ssvio
                 g2otypes.hpp compiled by oracle/Makefile from /root/reference, driven by
                 oracle/ref_harness.cpp (replays src/ssvio/backend.cpp:81-203).
  * PortOracle — oracle/_build/libssba_oracle.so: the plain-C restatement oracle/ssba_oracle.c.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  ssvio_b200 (the product) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libssba_ref.so")
PORT_SO = os.path.join(HERE, "_build", "libssba_oracle.so")
SHIM_SO = os.path.join(HERE, "_ref", "libssba_shim_test.so")
REFERENCE_ROOT = "/root/reference"

SSBA_MAX_ITER_RECORDS = 128


class IterRecord(C.Structure):
    _fields_ = [("chi2", C.c_double), ("lambda_", C.c_double), ("trials", C.c_int32),
                ("result", C.c_int32)]


class Report(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("last_result", C.c_int32),
                ("n_records", C.c_int32), ("cholesky_failures", C.c_int32),
                ("chi2_initial", C.c_double), ("chi2_robust", C.c_double),
                ("chi2_plain", C.c_double), ("lambda_", C.c_double),
                ("seconds_total", C.c_double), ("seconds_setup", C.c_double),
                ("iters", IterRecord * SSBA_MAX_ITER_RECORDS)]

    def trace(self):
        return [(self.iters[i].chi2, self.iters[i].lambda_, self.iters[i].trials)
                for i in range(self.n_records)]


class RefStats(C.Structure):
    _fields_ = [("t_residuals", C.c_double), ("t_quadratic_form", C.c_double),
                ("t_schur", C.c_double), ("t_linear_solver", C.c_double),
                ("t_linear_solution", C.c_double), ("t_update", C.c_double),
                ("t_initialize", C.c_double), ("t_graph_build", C.c_double),
                ("cholesky_nnz", C.c_int64), ("n_active_edges", C.c_int32),
                ("n_index_mapping", C.c_int32)]


def _p(a, typ):
    return a.ctypes.data_as(C.POINTER(typ)) if a is not None else None


def build(which=("port", "ref"), quiet=True):
    """Compile the checkers (oracle/Makefile). `ref` needs /root/reference and is skipped when
    it is absent (the GPU box uses the prebuilt .so that travelled with the snapshot)."""
    targets = []
    if "port" in which and os.path.exists(os.path.join(HERE, "ssba_oracle.c")):
        targets.append("port")
    if "ref" in which and os.path.isdir(REFERENCE_ROOT):
        targets.append("ref")
    if not targets:
        return
    subprocess.run(["make", "-C", HERE, "-j8"] + targets, check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


class _OracleBase:
    so_path = ""
    symbol = ""

    def __init__(self):
        if not os.path.exists(self.so_path):
            raise FileNotFoundError(f"{self.so_path} missing — run `make -C oracle`")
        self.lib = C.CDLL(self.so_path)

    @classmethod
    def available(cls):
        return os.path.exists(cls.so_path)


class RefOracle(_OracleBase):
    """The compiled reference (kind = "reference")."""
    so_path = REF_SO

    def __init__(self):
        super().__init__()
        f = self.lib.ssba_ref_optimize
        f.restype = C.c_int
        f.argtypes = [C.POINTER(C.c_double), C.c_int32, C.POINTER(C.c_double),
                      C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_uint8),
                      C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_uint8),
                      C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                      C.POINTER(C.c_uint8), C.POINTER(C.c_double), C.c_double,
                      C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_double,
                      C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double),
                      C.POINTER(Report), C.POINTER(RefStats), C.POINTER(C.c_int32),
                      C.POINTER(C.c_int64)]

    def optimize(self, g, iters=None, jacobian="numeric", trace=True, rounds=1,
                 outlier_threshold=5.891, want_state=True):
        iters = g.iters if iters is None else iters
        rep, st = Report(), RefStats()
        poses = np.empty_like(g.poses) if want_state else None
        points = np.empty_like(g.points) if want_state else None
        err = np.empty((g.n_edges, 2)) if want_state else None
        rd, no = C.c_int32(0), C.c_int64(0)
        rc = self.lib.ssba_ref_optimize(
            _p(g.K, C.c_double), g.ext.shape[0], _p(g.ext, C.c_double),
            g.n_poses, _p(g.poses, C.c_double), _p(g.pose_fixed, C.c_uint8),
            g.n_points, _p(g.points, C.c_double), _p(g.point_fixed, C.c_uint8),
            g.n_edges, _p(g.pose_idx, C.c_int32), _p(g.point_idx, C.c_int32),
            _p(g.cam_idx, C.c_uint8), _p(g.uv, C.c_double), float(g.huber_delta),
            int(iters), 1 if jacobian == "numeric" else 0, 1 if trace else 0, int(rounds),
            float(outlier_threshold), _p(poses, C.c_double), _p(points, C.c_double),
            _p(err, C.c_double), C.byref(rep), C.byref(st), C.byref(rd), C.byref(no))
        if rc != 0:
            raise RuntimeError(f"ssba_ref_optimize failed rc={rc}")
        return dict(report=rep, stats=st, poses=poses, points=points, errors=err,
                    rounds=rd.value, outliers=no.value)


def ref_pose_only(lib, batch, rounds=4, iters=10, chi2_threshold=5.991, pre_rounds=0):
    """FrontEnd::EstimateCurrentPose (pre_rounds = 0) / LoopClosing::OptimizeCurrentPose (pre_rounds = 1) through the
    compiled reference (ref_harness.cpp ssba_ref_pose_only_ex)."""
    nf, n = batch.n_frames, batch.xyz.shape[0]
    poses = np.empty((nf, 7)); flags = np.zeros(n, np.uint8); n_in = np.zeros(nf, np.int32); chi = np.zeros(nf)
    fp = np.ascontiguousarray(batch.feat_ptr, np.int32)
    rc = lib.ssba_ref_pose_only_ex(
        _p(np.ascontiguousarray(batch.K), C.c_double), nf, _p(fp, C.c_int32),
        _p(np.ascontiguousarray(batch.poses), C.c_double), _p(np.ascontiguousarray(batch.xyz), C.c_double),
        _p(np.ascontiguousarray(batch.uv), C.c_double), int(rounds), int(iters), C.c_double(chi2_threshold), C.c_int32(pre_rounds),
        _p(poses, C.c_double), _p(flags, C.c_uint8), _p(n_in, C.c_int32), _p(chi, C.c_double))
    if rc != 0:
        raise RuntimeError(f"ssba_ref_pose_only failed rc={rc}")
    return poses, flags, n_in, chi


def ref_pose_graph(lib, pg, iters=20):
    """LoopClosing::PoseGraphOptimization through the compiled reference (ssba_ref_pose_graph)."""
    rep = Report()
    out = np.empty_like(pg.poses)
    rc = lib.ssba_ref_pose_graph(
        pg.poses.shape[0], _p(np.ascontiguousarray(pg.poses), C.c_double), _p(np.ascontiguousarray(pg.fixed, np.uint8), C.c_uint8),
        len(pg.v0), _p(np.ascontiguousarray(pg.v0, np.int32), C.c_int32), _p(np.ascontiguousarray(pg.v1, np.int32), C.c_int32),
        _p(np.ascontiguousarray(pg.meas), C.c_double), int(iters), _p(out, C.c_double), C.byref(rep))
    if rc != 0:
        raise RuntimeError(f"ssba_ref_pose_graph failed rc={rc}")
    return out, rep


class PortOracle(_OracleBase):
    """The plain-C restatement (kind = "port")."""
    so_path = PORT_SO

    def __init__(self):
        super().__init__()
        f = self.lib.ssba_oracle_optimize
        f.restype = C.c_int
        f.argtypes = [C.POINTER(C.c_double), C.c_int32, C.POINTER(C.c_double),
                      C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_uint8),
                      C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_uint8),
                      C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                      C.POINTER(C.c_uint8), C.POINTER(C.c_double), C.c_double,
                      C.c_int32, C.c_int32,
                      C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double),
                      C.POINTER(Report)]

    def set_tau(self, tau: float = 1e-5):
        """_tau of the Levenberg initialisation (levenberg.cpp:44-51); 1e-5 is g2o's default."""
        self.lib.ssba_oracle_set_tau.argtypes = [C.c_double]
        self.lib.ssba_oracle_set_tau.restype = None
        self.lib.ssba_oracle_set_tau(float(tau))

    def optimize(self, g, iters=None, jacobian="numeric"):
        iters = g.iters if iters is None else iters
        rep = Report()
        poses = np.empty_like(g.poses)
        points = np.empty_like(g.points)
        err = np.empty((g.n_edges, 2))
        rc = self.lib.ssba_oracle_optimize(
            _p(g.K, C.c_double), g.ext.shape[0], _p(g.ext, C.c_double),
            g.n_poses, _p(g.poses, C.c_double), _p(g.pose_fixed, C.c_uint8),
            g.n_points, _p(g.points, C.c_double), _p(g.point_fixed, C.c_uint8),
            g.n_edges, _p(g.pose_idx, C.c_int32), _p(g.point_idx, C.c_int32),
            _p(g.cam_idx, C.c_uint8), _p(g.uv, C.c_double), float(g.huber_delta),
            int(iters), 1 if jacobian == "numeric" else 0,
            _p(poses, C.c_double), _p(points, C.c_double), _p(err, C.c_double), C.byref(rep))
        if rc != 0:
            raise RuntimeError(f"ssba_oracle_optimize failed rc={rc}")
        return dict(report=rep, poses=poses, points=points, errors=err)


class ShimHarness(_OracleBase):
    """Drop-in proof (oracle/shim_harness.cpp): the reference's graph construction and g2o
    SparseOptimizer with include/ssba_g2o_shim.hpp as the optimisation algorithm, i.e. the CUDA
    path reached exactly the way ssvio's backend.cpp would reach it.  Needs a GPU to run."""
    so_path = SHIM_SO

    def __init__(self):
        super().__init__()
        f = self.lib.ssba_shim_optimize
        f.restype = C.c_int
        f.argtypes = [C.POINTER(C.c_double), C.c_int32, C.POINTER(C.c_double),
                      C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_uint8),
                      C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_uint8),
                      C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                      C.POINTER(C.c_uint8), C.POINTER(C.c_double), C.c_double, C.c_int32,
                      C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double),
                      C.POINTER(Report)]

    def set_write_back_every_iteration(self, on: bool):
        """False: one writeBack() after optimize() (all backend.cpp needs) instead of three read-backs per iteration."""
        self.lib.ssba_shim_set_write_back_every_iteration.argtypes = [C.c_int32]
        self.lib.ssba_shim_set_write_back_every_iteration.restype = None
        self.lib.ssba_shim_set_write_back_every_iteration(1 if on else 0)

    def optimize(self, g, iters=None):
        iters = g.iters if iters is None else iters
        rep = Report()
        poses = np.empty_like(g.poses)
        points = np.empty_like(g.points)
        chi2 = np.empty(g.n_edges)
        rc = self.lib.ssba_shim_optimize(
            _p(g.K, C.c_double), g.ext.shape[0], _p(g.ext, C.c_double),
            g.n_poses, _p(g.poses, C.c_double), _p(g.pose_fixed, C.c_uint8),
            g.n_points, _p(g.points, C.c_double), _p(g.point_fixed, C.c_uint8),
            g.n_edges, _p(g.pose_idx, C.c_int32), _p(g.point_idx, C.c_int32),
            _p(g.cam_idx, C.c_uint8), _p(g.uv, C.c_double), float(g.huber_delta), int(iters),
            _p(poses, C.c_double), _p(points, C.c_double), _p(chi2, C.c_double), C.byref(rep))
        if rc != 0:
            raise RuntimeError(f"ssba_shim_optimize failed rc={rc}")
        return dict(report=rep, poses=poses, points=points, edge_chi2=chi2)
