"""Thin ctypes binding of libssba (include/ssba.h) for tests and bench.py.

The product is the C-ABI shared library; this module only marshals numpy arrays into it.  The
method names follow the reference's call sequence in Backend::OptimizeActiveMap()
(src/ssvio/backend.cpp:81-203): add vertices / edges, initializeOptimization(), optimize(N),
then read estimates and per-edge chi2.

There is no fallback: if ssvio_b200/lib/libssba.so is missing or no CUDA device is usable, this
raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# SSBA_LIB: developer override (e.g. a -DSSBA_SOLVER_TRACE build); the product path is the in-tree library
LIB_PATH = os.environ.get("SSBA_LIB") or os.path.join(HERE, "lib", "libssba.so")

SSBA_MAX_ITER_RECORDS = 128
SSBA_NCCL_ID_BYTES = 128

STATUS = {0: "OK", 1: "INVALID_ARG", 2: "CUDA", 3: "NO_DEVICE", 4: "STATE", 5: "NCCL", 6: "EMPTY",
          7: "ALLOC"}
SSBA_ERR_EMPTY = 6
SOLVER_OK, SOLVER_TERMINATE, SOLVER_FAIL = 1, 2, -1

# every symbol include/ssba.h declares (tests check the library exports all of them)
ABI_SYMBOLS = [
    "ssba_default_options", "ssba_create", "ssba_destroy", "ssba_last_error",
    "ssba_nccl_unique_id", "ssba_set_cameras", "ssba_set_poses", "ssba_set_points",
    "ssba_set_edges", "ssba_initialize", "ssba_optimize", "ssba_step", "ssba_reset_state", "ssba_drop_structure",
    "ssba_get_poses", "ssba_get_points", "ssba_get_edge_errors", "ssba_chi2",
    "ssba_count_outliers", "ssba_get_outlier_mask", "ssba_request_stop", "ssba_clear_stop", "ssba_optimize_rounds", "ssba_plan_shards", "ssba_set_profiling", "ssba_profile_get",
    "ssba_profile_reset", "ssba_get_problem_info", "ssba_pose_only_optimize", "ssba_pose_only_optimize_loop",
    "ssba_pose_graph_optimize",
    "ssba_version",
]


class Options(C.Structure):
    _fields_ = [("tau", C.c_double), ("good_step_lower_scale", C.c_double),
                ("good_step_upper_scale", C.c_double), ("user_lambda_init", C.c_double),
                ("max_trials_after_failure", C.c_int32), ("jacobian_mode", C.c_int32),
                ("device_id", C.c_int32), ("profile", C.c_int32), ("stream", C.c_void_p),
                ("rank", C.c_int32), ("world_size", C.c_int32),
                ("nccl_id", C.c_uint8 * SSBA_NCCL_ID_BYTES), ("presharded", C.c_int32), ("reserved", C.c_int32 * 7)]


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


class Profile(C.Structure):
    _fields_ = [("ms_linearize", C.c_double), ("ms_schur", C.c_double),
                ("ms_reduced_solve", C.c_double), ("ms_update_chi2", C.c_double),
                ("ms_allreduce", C.c_double), ("n_linearize", C.c_int64), ("n_schur", C.c_int64),
                ("n_reduced_solve", C.c_int64), ("n_update_chi2", C.c_int64),
                ("n_allreduce", C.c_int64), ("kernel_launches", C.c_int64),
                ("levenberg_iterations", C.c_int64), ("outer_iterations", C.c_int64), ("cholesky_nnz", C.c_int64),
                ("hessian_pose_dimension", C.c_int32), ("hessian_landmark_dimension", C.c_int32),
                ("ms_symbolic_decomposition", C.c_double), ("ms_numeric_decomposition", C.c_double),
                ("ms_structure_build", C.c_double)]


class ProblemInfo(C.Structure):
    _fields_ = [("n_free_poses", C.c_int32), ("n_free_points", C.c_int32),
                ("n_active_edges", C.c_int32), ("n_pairs", C.c_int32),
                ("n_schur_blocks", C.c_int32), ("n_factor_blocks", C.c_int32),
                ("device_bytes", C.c_int64), ("solve_cluster", C.c_int32), ("peer_exchange", C.c_int32),
                ("solver_kind", C.c_int32), ("solver_steps", C.c_int32), ("solver_top_cols", C.c_int32),
                ("solver_smem_bytes", C.c_int32), ("n_structure_builds", C.c_int64), ("n_structure_reuses", C.c_int64)]


class SsbaError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"libssba: {STATUS.get(status, status)}: {msg}")
        self.status = status


_lib = None


def load_library():
    """Load libssba.so; fails loudly when it was not built (no silent fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} not built — run `python -m ssvio_b200.build` "
                          "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    H = C.c_void_p
    dp, ip, bp = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_uint8)
    lib.ssba_default_options.argtypes = [C.POINTER(Options)]
    lib.ssba_default_options.restype = None
    lib.ssba_create.argtypes = [C.POINTER(Options), C.POINTER(H)]
    lib.ssba_destroy.argtypes = [H]
    lib.ssba_destroy.restype = None
    lib.ssba_last_error.argtypes = [H]
    lib.ssba_last_error.restype = C.c_char_p
    lib.ssba_nccl_unique_id.argtypes = [bp]
    lib.ssba_set_cameras.argtypes = [H, dp, C.c_int32, dp]
    lib.ssba_set_poses.argtypes = [H, C.c_int32, dp, bp]
    lib.ssba_set_points.argtypes = [H, C.c_int32, dp, bp]
    lib.ssba_set_edges.argtypes = [H, C.c_int32, ip, ip, bp, dp, dp, dp, C.c_double]
    lib.ssba_initialize.argtypes = [H]
    lib.ssba_optimize.argtypes = [H, C.c_int32, C.POINTER(Report)]
    lib.ssba_step.argtypes = [H, C.c_int32, C.POINTER(C.c_int32), C.POINTER(IterRecord)]
    lib.ssba_reset_state.argtypes = [H]
    lib.ssba_drop_structure.argtypes = [H]
    lib.ssba_get_poses.argtypes = [H, dp]
    lib.ssba_get_points.argtypes = [H, dp]
    lib.ssba_get_edge_errors.argtypes = [H, dp]
    lib.ssba_chi2.argtypes = [H, dp, dp]
    lib.ssba_count_outliers.argtypes = [H, C.c_double, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    lib.ssba_get_outlier_mask.argtypes = [H, C.c_double, bp, C.POINTER(C.c_int64)]
    lib.ssba_request_stop.argtypes = [H]
    lib.ssba_clear_stop.argtypes = [H]
    lib.ssba_optimize_rounds.argtypes = [H, C.c_int32, C.c_int32, C.c_double, C.c_double, C.POINTER(C.c_int32),
                                         C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(Report)]
    lib.ssba_plan_shards.argtypes = [C.c_int32, bp, C.c_int32, bp, C.c_int32, ip, ip, C.c_int32, ip]
    lib.ssba_pose_only_optimize.argtypes = [H, dp, C.c_int32, ip, dp, dp, dp, C.c_int32, C.c_int32, C.c_double,
                                            dp, bp, ip, dp]
    lib.ssba_pose_only_optimize_loop.argtypes = lib.ssba_pose_only_optimize.argtypes
    lib.ssba_pose_graph_optimize.argtypes = [H, C.c_int32, dp, bp, C.c_int32, ip, ip, dp, C.c_int32, dp, C.POINTER(Report)]
    lib.ssba_set_profiling.argtypes = [H, C.c_int32]
    lib.ssba_profile_get.argtypes = [H, C.POINTER(Profile)]
    lib.ssba_profile_reset.argtypes = [H]
    lib.ssba_get_problem_info.argtypes = [H, C.POINTER(ProblemInfo)]
    lib.ssba_version.restype = C.c_int32
    for name in ABI_SYMBOLS:
        f = getattr(lib, name)
        if f.restype is C.c_int and name not in ("ssba_version",):
            f.restype = C.c_int
    _lib = lib
    return lib


def _p(a, typ):
    return a.ctypes.data_as(C.POINTER(typ)) if a is not None else None


def _c(a, dtype):
    return None if a is None else np.ascontiguousarray(a, dtype=dtype)


def nccl_unique_id() -> bytes:
    lib = load_library()
    buf = (C.c_uint8 * SSBA_NCCL_ID_BYTES)()
    st = lib.ssba_nccl_unique_id(buf)
    if st != 0:
        raise SsbaError(st, lib.ssba_last_error(None).decode())
    return bytes(buf)


def plan_shards(g, world_size: int) -> np.ndarray:
    """Host-only: owner rank of every landmark of graph `g` when sharded over `world_size`."""
    lib = load_library()
    owner = np.empty(g.n_points, dtype=np.int32)
    pf, lf = _c(g.pose_fixed, np.uint8), _c(g.point_fixed, np.uint8)
    pi, li = _c(g.pose_idx, np.int32), _c(g.point_idx, np.int32)
    st = lib.ssba_plan_shards(g.n_poses, _p(pf, C.c_uint8), g.n_points, _p(lf, C.c_uint8), g.n_edges,
                              _p(pi, C.c_int32), _p(li, C.c_int32), int(world_size), _p(owner, C.c_int32))
    if st != 0:
        raise SsbaError(st, lib.ssba_last_error(None).decode())
    return owner


class BundleAdjuster:
    """One g2o::SparseOptimizer + Levenberg + BlockSolver_6_3 + CSparse stack
    (backend.cpp:81-86), resident on one B200."""

    def __init__(self, *, jacobian="analytic", device_id=-1, stream=None, profile=False,
                 rank=0, world_size=1, nccl_id: bytes | None = None, user_lambda_init=0.0,
                 max_trials_after_failure=10, tau=None, presharded=False):
        self.lib = load_library()
        opt = Options()
        self.lib.ssba_default_options(C.byref(opt))
        opt.jacobian_mode = 1 if jacobian == "numeric" else 0
        opt.device_id = device_id
        opt.profile = 1 if profile else 0
        opt.stream = stream
        opt.rank, opt.world_size = rank, world_size
        opt.presharded = 1 if presharded else 0
        opt.user_lambda_init = user_lambda_init
        opt.max_trials_after_failure = max_trials_after_failure
        if tau is not None:
            opt.tau = tau  # _tau of the Levenberg initialisation (levenberg.cpp:44-51)
        if nccl_id is not None:
            C.memmove(opt.nccl_id, nccl_id, SSBA_NCCL_ID_BYTES)
        self._h = C.c_void_p()
        st = self.lib.ssba_create(C.byref(opt), C.byref(self._h))
        if st != 0:
            raise SsbaError(st, self.lib.ssba_last_error(None).decode())
        self._n_poses = self._n_points = self._n_edges = 0

    # -- life cycle
    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self.lib.ssba_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, st):
        if st != 0:
            raise SsbaError(st, self.lib.ssba_last_error(self._h).decode())

    # -- graph construction (backend.cpp:88-168)
    def set_cameras(self, K, ext):
        K = _c(np.asarray(K).reshape(-1), np.float64)
        ext = _c(np.asarray(ext).reshape(-1, 7), np.float64)
        self._check(self.lib.ssba_set_cameras(self._h, _p(K, C.c_double), ext.shape[0],
                                              _p(ext, C.c_double)))

    def set_poses(self, poses, fixed=None):
        poses = _c(poses, np.float64)
        fixed = _c(fixed, np.uint8)
        self._n_poses = poses.shape[0]
        self._check(self.lib.ssba_set_poses(self._h, poses.shape[0], _p(poses, C.c_double),
                                            _p(fixed, C.c_uint8)))

    def set_points(self, points, fixed=None):
        points = _c(points, np.float64)
        fixed = _c(fixed, np.uint8)
        self._n_points = points.shape[0]
        self._check(self.lib.ssba_set_points(self._h, points.shape[0], _p(points, C.c_double),
                                             _p(fixed, C.c_uint8)))

    def set_edges(self, pose_idx, point_idx, cam_idx, uv, info=None, huber_delta=None,
                  huber_delta_all=0.0):
        pose_idx = _c(pose_idx, np.int32)
        point_idx = _c(point_idx, np.int32)
        cam_idx = _c(cam_idx, np.uint8)
        uv = _c(uv, np.float64)
        info = _c(info, np.float64)
        huber_delta = _c(huber_delta, np.float64)
        self._n_edges = pose_idx.shape[0]
        self._check(self.lib.ssba_set_edges(
            self._h, pose_idx.shape[0], _p(pose_idx, C.c_int32), _p(point_idx, C.c_int32),
            _p(cam_idx, C.c_uint8), _p(uv, C.c_double), _p(info, C.c_double),
            _p(huber_delta, C.c_double), float(huber_delta_all)))

    def set_graph(self, g):
        """Upload a ssvio_b200.synth.Graph (the same arrays the oracles get)."""
        self.set_cameras(g.K, g.ext)
        self.set_poses(g.poses, g.pose_fixed)
        self.set_points(g.points, g.point_fixed)
        self.set_edges(g.pose_idx, g.point_idx, g.cam_idx, g.uv, huber_delta_all=g.huber_delta)

    # -- optimisation (backend.cpp:177-178)
    def initialize_optimization(self):
        self._check(self.lib.ssba_initialize(self._h))

    def optimize(self, iterations) -> Report:
        rep = Report()
        st = self.lib.ssba_optimize(self._h, int(iterations), C.byref(rep))
        if st == SSBA_ERR_EMPTY:
            return rep  # iterations == -1, like SparseOptimizer::optimize()
        self._check(st)
        return rep

    def optimize_nowait_report(self, iterations):
        """optimize() without the final chi2 read-out (bench inner loop)."""
        self._check(self.lib.ssba_optimize(self._h, int(iterations), None))

    def optimize_rounds(self, max_rounds=5, iters_per_round=10, chi2_threshold=5.891, inlier_ratio=0.7):
        """The round loop of backend.cpp:175-203; returns (rounds, n_outliers, n_inliers, report)."""
        rd, no, ni, rep = C.c_int32(0), C.c_int64(0), C.c_int64(0), Report()
        self._check(self.lib.ssba_optimize_rounds(self._h, int(max_rounds), int(iters_per_round), float(chi2_threshold),
                                                  float(inlier_ratio), C.byref(rd), C.byref(no), C.byref(ni), C.byref(rep)))
        return rd.value, no.value, ni.value, rep

    def step(self, iteration):
        res = C.c_int32(0)
        rec = IterRecord()
        self._check(self.lib.ssba_step(self._h, int(iteration), C.byref(res), C.byref(rec)))
        return res.value, rec

    def reset_state(self):
        self._check(self.lib.ssba_reset_state(self._h))

    def drop_structure(self):
        """Force the next initialize to rebuild the structure even for an unchanged topology."""
        self._check(self.lib.ssba_drop_structure(self._h))

    # -- results (backend.cpp:180-244)
    def poses(self):
        out = np.empty((self._n_poses, 7))
        self._check(self.lib.ssba_get_poses(self._h, _p(out, C.c_double)))
        return out

    def points(self):
        out = np.empty((self._n_points, 3))
        self._check(self.lib.ssba_get_points(self._h, _p(out, C.c_double)))
        return out

    def edge_errors(self):
        out = np.empty((self._n_edges, 2))
        self._check(self.lib.ssba_get_edge_errors(self._h, _p(out, C.c_double)))
        return out

    def chi2(self):
        a, b = C.c_double(0), C.c_double(0)
        self._check(self.lib.ssba_chi2(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def count_outliers(self, threshold=5.891):
        a, b = C.c_int64(0), C.c_int64(0)
        self._check(self.lib.ssba_count_outliers(self._h, float(threshold), C.byref(a), C.byref(b)))
        return a.value, b.value

    def outlier_mask(self, threshold=5.891):
        """uint8 per edge (caller's order): plain chi2 > threshold (backend.cpp:205-227)."""
        out = np.zeros(self._n_edges, np.uint8)
        n = C.c_int64(0)
        self._check(self.lib.ssba_get_outlier_mask(self._h, float(threshold), _p(out, C.c_uint8), C.byref(n)))
        return out, n.value

    def request_stop(self):
        self._check(self.lib.ssba_request_stop(self._h))

    def clear_stop(self):
        self._check(self.lib.ssba_clear_stop(self._h))

    def profile(self) -> Profile:
        p = Profile()
        self._check(self.lib.ssba_profile_get(self._h, C.byref(p)))
        return p

    def profile_reset(self):
        self._check(self.lib.ssba_profile_reset(self._h))

    # -- pose-only LM of the front-end (frontend.cpp:184-260), batched over frames
    def pose_only_optimize(self, batch, rounds=4, iters=10, chi2_threshold=5.991, loop_closing=False):
        """batch: ssvio_b200.synth.PoseOnlyBatch.  Returns (poses, outlier flags, inliers per frame,
        robust chi2 per frame).  loop_closing: the schedule of LoopClosing::OptimizeCurrentPose
        (loopclosing.cpp:245-351: one unclassified optimize() before the rounds)."""
        nf, n = batch.n_frames, int(batch.xyz.shape[0])
        K, fp = _c(batch.K, np.float64), _c(batch.feat_ptr, np.int32)
        pin, xyz, uv = _c(batch.poses, np.float64), _c(batch.xyz, np.float64), _c(batch.uv, np.float64)
        poses, chi = np.empty((nf, 7)), np.zeros(nf)
        flags, n_in = np.zeros(n, np.uint8), np.zeros(nf, np.int32)
        fn = self.lib.ssba_pose_only_optimize_loop if loop_closing else self.lib.ssba_pose_only_optimize
        self._check(fn(
            self._h, _p(K, C.c_double), nf, _p(fp, C.c_int32), _p(pin, C.c_double), _p(xyz, C.c_double),
            _p(uv, C.c_double), int(rounds), int(iters), float(chi2_threshold), _p(poses, C.c_double),
            _p(flags, C.c_uint8), _p(n_in, C.c_int32), _p(chi, C.c_double)))
        return poses, flags, n_in, chi

    # -- pose-graph optimisation of the loop closer (loopclosing.cpp:458-532)
    def pose_graph_optimize(self, pg, iters=20):
        """pg: ssvio_b200.synth.PoseGraph.  Returns (poses, Report)."""
        poses_in, fixed = _c(pg.poses, np.float64), _c(pg.fixed, np.uint8)
        v0, v1, meas = _c(pg.v0, np.int32), _c(pg.v1, np.int32), _c(pg.meas, np.float64)
        out, rep = np.empty_like(poses_in), Report()
        st = self.lib.ssba_pose_graph_optimize(self._h, poses_in.shape[0], _p(poses_in, C.c_double), _p(fixed, C.c_uint8),
                                               len(v0), _p(v0, C.c_int32), _p(v1, C.c_int32), _p(meas, C.c_double), int(iters),
                                               _p(out, C.c_double), C.byref(rep))
        if st == SSBA_ERR_EMPTY:
            return out, rep
        self._check(st)
        return out, rep

    def set_profiling(self, on: bool):
        self._check(self.lib.ssba_set_profiling(self._h, 1 if on else 0))

    def problem_info(self) -> ProblemInfo:
        p = ProblemInfo()
        self._check(self.lib.ssba_get_problem_info(self._h, C.byref(p)))
        return p
