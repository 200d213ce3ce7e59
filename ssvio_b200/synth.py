"""KITTI-shaped synthetic local-BA graphs — the workload of BASELINE.json.

The generator is the single source of inputs for the reference oracle, the C restatement and
the CUDA path: all three are fed the SAME flat arrays, in the layout of include/ssba.h.  It
draws only uniform doubles from numpy's PCG64 (a stream numpy guarantees stable) and builds
normals by Box-Muller, so a (config, seed) pair names one graph on every machine.

Shapes follow the reference's data model: one pose per key-frame (T_cw, ssvio::VertexPose,
include/ssvio/g2otypes.hpp:28-46), one 3-D landmark per map-point (VertexXYZ, :49-65), one
2-D EdgeProjection per (key-frame, landmark, camera) (:112-131), intrinsics and the
right-camera extrinsic of config/kitti_00.yaml:3-6,26 / src/ssvio/system.cpp:69-71, Huber
delta 5.891 (src/ssvio/backend.cpp:109,163).
"""
from __future__ import annotations

import dataclasses
import numpy as np

FX = FY = 718.856
CX, CY = 607.1928, 185.2157
BF = 386.1448
HUBER_DELTA = 5.891


@dataclasses.dataclass
class Graph:
    """Flat local-BA problem in the C-ABI layout (include/ssba.h)."""

    K: np.ndarray            # (9,)  row-major 3x3
    ext: np.ndarray          # (n_cams, 7) qx qy qz qw tx ty tz
    poses: np.ndarray        # (NK, 7) initial T_cw
    pose_fixed: np.ndarray   # (NK,) uint8
    points: np.ndarray       # (NP, 3) initial landmarks
    point_fixed: np.ndarray  # (NP,) uint8
    pose_idx: np.ndarray     # (E,) int32
    point_idx: np.ndarray    # (E,) int32
    cam_idx: np.ndarray      # (E,) uint8
    uv: np.ndarray           # (E, 2)
    huber_delta: float = HUBER_DELTA
    name: str = ""
    iters: int = 10

    @property
    def n_poses(self) -> int:
        return int(self.poses.shape[0])

    @property
    def n_points(self) -> int:
        return int(self.points.shape[0])

    @property
    def n_edges(self) -> int:
        return int(self.pose_idx.shape[0])

    def input_bytes(self) -> int:
        return int(sum(a.nbytes for a in (self.K, self.ext, self.poses, self.pose_fixed,
                                          self.points, self.point_fixed, self.pose_idx,
                                          self.point_idx, self.cam_idx, self.uv)))

    def output_bytes(self) -> int:
        return int(self.poses.nbytes + self.points.nbytes)

    def shard(self, rank: int, world: int) -> "Graph":
        """The part of the graph rank `rank` of `world` is handed with ssba_options.presharded: all edges of the
        landmarks it owns (contiguous ranges of point rows holding about 1 / world of the edges each); poses, fixed
        flags and the point array (indexed globally) are the same on every rank."""
        deg = np.bincount(self.point_idx, minlength=self.n_points)
        cum = np.cumsum(deg)
        bounds = np.searchsorted(cum, [cum[-1] * r / world for r in range(1, world)], side="left") + 1
        lo = 0 if rank == 0 else int(bounds[rank - 1])
        hi = self.n_points if rank == world - 1 else int(bounds[rank])
        sel = np.nonzero((self.point_idx >= lo) & (self.point_idx < hi))[0]
        return Graph(K=self.K, ext=self.ext, poses=self.poses, pose_fixed=self.pose_fixed, points=self.points,
                     point_fixed=self.point_fixed, pose_idx=np.ascontiguousarray(self.pose_idx[sel]),
                     point_idx=np.ascontiguousarray(self.point_idx[sel]), cam_idx=np.ascontiguousarray(self.cam_idx[sel]),
                     uv=np.ascontiguousarray(self.uv[sel]), huber_delta=self.huber_delta, iters=self.iters)


# --------------------------------------------------------------------------- SE(3) helpers
# Sophus conventions (thirdparty/sophus/sophus/so3.hpp:593-622, se3.hpp:763-784): tangent is
# (upsilon, omega); quaternions are stored (x, y, z, w).

def _quat_mul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    ax, ay, az, aw = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    bx, by, bz, bw = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    q = np.stack([aw * bx + ax * bw + ay * bz - az * by,
                  aw * by + ay * bw + az * bx - ax * bz,
                  aw * bz + az * bw + ax * by - ay * bx,
                  aw * bw - ax * bx - ay * by - az * bz], axis=-1)
    return q / np.linalg.norm(q, axis=-1, keepdims=True)


def _quat_rot(q: np.ndarray, p: np.ndarray) -> np.ndarray:
    v, w = q[..., :3], q[..., 3:4]
    uv = 2.0 * np.cross(v, p)
    return p + w * uv + np.cross(v, uv)


def se3_exp(a: np.ndarray) -> np.ndarray:
    """(N, 6) tangents -> (N, 7) poses (qx qy qz qw tx ty tz)."""
    a = np.atleast_2d(np.asarray(a, dtype=np.float64))
    ups, om = a[:, :3], a[:, 3:]
    th2 = np.sum(om * om, axis=1)
    th = np.sqrt(th2)
    small = th < 1e-10
    th_s = np.where(small, 1.0, th)
    imag = np.where(small, 0.5 - th2 / 48.0, np.sin(0.5 * th_s) / th_s)
    real = np.where(small, 1.0 - th2 / 8.0, np.cos(0.5 * th_s))
    q = np.concatenate([imag[:, None] * om, real[:, None]], axis=1)
    # V = I + (1-cos)/th^2 * hat + (th - sin)/th^3 * hat^2
    c1 = np.where(small, 0.5, (1.0 - np.cos(th_s)) / (th_s * th_s))
    c2 = np.where(small, 1.0 / 6.0, (th_s - np.sin(th_s)) / (th_s ** 3))
    wxu = np.cross(om, ups)
    t = ups + c1[:, None] * wxu + c2[:, None] * np.cross(om, wxu)
    return np.concatenate([q, t], axis=1)


def se3_mul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    q = _quat_mul(a[..., :4], b[..., :4])
    t = a[..., 4:] + _quat_rot(a[..., :4], b[..., 4:])
    return np.concatenate([q, t], axis=-1)


def se3_inv(a: np.ndarray) -> np.ndarray:
    qi = a[..., :4] * np.array([-1.0, -1.0, -1.0, 1.0])
    return np.concatenate([qi, -_quat_rot(qi, a[..., 4:])], axis=-1)


def se3_act(a: np.ndarray, p: np.ndarray) -> np.ndarray:
    return _quat_rot(a[..., :4], p) + a[..., 4:]


# --------------------------------------------------------------------------- RNG

class _Rng:
    def __init__(self, seed: int):
        self._g = np.random.Generator(np.random.PCG64(seed))

    def uniform(self, lo, hi, size):
        return lo + (hi - lo) * self._g.random(size)

    def normal(self, size):
        n = int(np.prod(size))
        m = (n + 1) // 2
        u1 = 1.0 - self._g.random(m)  # (0, 1]
        u2 = self._g.random(m)
        r = np.sqrt(-2.0 * np.log(u1))
        z = np.concatenate([r * np.cos(2 * np.pi * u2), r * np.sin(2 * np.pi * u2)])[:n]
        return z.reshape(size)

    def integers(self, lo, hi_inclusive, size):
        u = self._g.random(size)
        return np.minimum((lo + np.floor(u * (hi_inclusive - lo + 1))).astype(np.int64),
                          hi_inclusive)


# --------------------------------------------------------------------------- generator

def make_graph(n_kf: int, n_points: int, window: int, *, seed: int = 42,
               fix_first_pose: bool = False, pixel_sigma: float = 0.7,
               outlier_frac: float = 0.02, outlier_sigma: float = 30.0,
               n_fixed_points: int = 0, name: str = "", iters: int = 10) -> Graph:
    """One sliding-window graph: every landmark is seen by `window` consecutive key-frames in
    both cameras -> n_points * window * 2 edges, added landmark-major, key-frame-minor, left
    then right."""
    if window > n_kf:
        raise ValueError("window larger than the number of key-frames")
    rng = _Rng(seed)
    K = np.array([FX, 0, CX, 0, FY, CY, 0, 0, 1], dtype=np.float64)
    ext = np.array([[0, 0, 0, 1, 0, 0, 0],
                    [0, 0, 0, 1, -BF / FX, 0, 0]], dtype=np.float64)

    # ground-truth trajectory: ~1 m forward per key-frame (T_cw, so -z), small jitter
    n = rng.normal((n_kf, 6))
    tang = np.stack([0.02 * n[:, 0], 0.01 * n[:, 1], -1.0 * np.arange(n_kf), 0.002 * n[:, 3],
                     0.002 * n[:, 4], 0.002 * n[:, 5]], axis=1)
    T_gt = se3_exp(tang)
    d = np.concatenate([0.05 * rng.normal((n_kf, 3)), 0.005 * rng.normal((n_kf, 3))], axis=1)
    if fix_first_pose:
        d[0] = 0.0
    poses = se3_mul(se3_exp(d), T_gt)
    pose_fixed = np.zeros(n_kf, dtype=np.uint8)
    if fix_first_pose:
        pose_fixed[0] = 1

    k0 = rng.integers(0, n_kf - window, n_points)
    pc = np.stack([rng.uniform(-15, 15, n_points), rng.uniform(-3, 3, n_points),
                   rng.uniform(8, 48, n_points)], axis=1)
    pw = se3_act(se3_inv(T_gt[k0]), pc)
    points = pw + 0.2 * rng.normal((n_points, 3))
    point_fixed = np.zeros(n_points, dtype=np.uint8)
    if n_fixed_points > 0:
        # landmarks "anchored to a key-frame that left the window" (backend.cpp:125-130):
        # fixed at their (noisy) position, spread evenly over the landmark list
        point_fixed[np.linspace(0, n_points - 1, n_fixed_points).astype(np.int64)] = 1

    # edges: landmark-major, key-frame-minor, left then right
    lm = np.repeat(np.arange(n_points), window * 2)
    kf = (k0[:, None, None] + np.arange(window)[None, :, None] + np.zeros((1, 1, 2), np.int64))
    kf = kf.reshape(-1)
    cam = np.tile(np.array([0, 1], dtype=np.uint8), n_points * window)
    pb = se3_act(T_gt[kf], pw[lm])
    pcam = se3_act(ext[cam], pb)
    keep = pcam[:, 2] > 0.5
    u = FX * pcam[:, 0] / pcam[:, 2] + CX
    v = FY * pcam[:, 1] / pcam[:, 2] + CY
    uv = np.stack([u, v], axis=1) + pixel_sigma * rng.normal((lm.shape[0], 2))
    is_out = rng.uniform(0, 1, lm.shape[0]) < outlier_frac
    uv = uv + is_out[:, None] * (outlier_sigma * rng.normal((lm.shape[0], 2)))

    return Graph(K=K, ext=ext, poses=np.ascontiguousarray(poses), pose_fixed=pose_fixed,
                 points=np.ascontiguousarray(points), point_fixed=point_fixed,
                 pose_idx=np.ascontiguousarray(kf[keep].astype(np.int32)),
                 point_idx=np.ascontiguousarray(lm[keep].astype(np.int32)),
                 cam_idx=np.ascontiguousarray(cam[keep]),
                 uv=np.ascontiguousarray(uv[keep]), name=name, iters=iters)


# BASELINE.json configs (NK, NP, window, LM iterations, fix KF0)
CONFIGS = {
    "cfg1": dict(n_kf=10, n_points=500, window=3, iters=5),
    "cfg2": dict(n_kf=50, n_points=5000, window=4, iters=10),
    "cfg3": dict(n_kf=100, n_points=20000, window=5, iters=10),
    "cfg5": dict(n_kf=500, n_points=100000, window=5, iters=10, fix_first_pose=True),
    # small shapes for fast parity tests
    "tiny": dict(n_kf=4, n_points=40, window=3, iters=5),
    "small": dict(n_kf=8, n_points=200, window=4, iters=8),
}


def make_config(name: str, seed: int = 42, **overrides) -> Graph:
    kw = dict(CONFIGS[name])
    kw.update(overrides)
    iters = kw.pop("iters", 10)
    return make_graph(seed=seed, name=name, iters=iters, **kw)


# --------------------------------------------------------------------------- pose-only frames

@dataclasses.dataclass
class PoseOnlyBatch:
    """A batch of front-end pose estimations (src/ssvio/frontend.cpp:184-260): per frame an initial
    T_cw and its features = (map-point position, measured pixel in the left camera)."""

    K: np.ndarray         # (9,)
    feat_ptr: np.ndarray  # (F + 1,) int32
    poses: np.ndarray     # (F, 7) initial T_cw
    xyz: np.ndarray       # (N, 3)
    uv: np.ndarray        # (N, 2)

    @property
    def n_frames(self) -> int:
        return int(self.poses.shape[0])


def make_pose_only(n_frames: int, n_features: int, *, seed: int = 42, pixel_sigma: float = 0.7,
                   outlier_frac: float = 0.1, outlier_sigma: float = 25.0, pose_sigma_t: float = 0.08,
                   pose_sigma_r: float = 0.01) -> PoseOnlyBatch:
    """Frames along the KITTI-shaped trajectory of make_graph; every frame sees `n_features`
    (+-25 %) map points 8..48 m ahead; 10 % of the measurements are gross outliers, which the
    reference's four rounds have to find; the initial pose is the truth perturbed like a
    constant-velocity prediction would be."""
    rng = _Rng(seed)
    K = np.array([FX, 0, CX, 0, FY, CY, 0, 0, 1], dtype=np.float64)
    counts = np.maximum(8, (n_features * rng.uniform(0.75, 1.25, n_frames)).astype(np.int64))
    feat_ptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    n = int(feat_ptr[-1])
    idx = np.arange(n_frames, dtype=np.float64)
    gt = se3_exp(np.stack([0.02 * idx, 0.01 * idx, -1.0 * idx, 0.002 * idx, 0.002 * idx, 0.002 * idx], axis=1))
    frame_of = np.repeat(np.arange(n_frames), counts)
    pc = np.stack([rng.uniform(-15, 15, n), rng.uniform(-3, 3, n), rng.uniform(8, 48, n)], axis=1)
    xyz = se3_act(se3_inv(gt)[frame_of], pc)
    uv = np.stack([FX * pc[:, 0] / pc[:, 2] + CX, FY * pc[:, 1] / pc[:, 2] + CY], axis=1)
    uv = uv + pixel_sigma * rng.normal((n, 2))
    bad = rng.uniform(0, 1, n) < outlier_frac
    uv = uv + bad[:, None] * outlier_sigma * rng.normal((n, 2))
    d = np.concatenate([pose_sigma_t * rng.normal((n_frames, 3)), pose_sigma_r * rng.normal((n_frames, 3))], axis=1)
    poses = se3_mul(se3_exp(d), gt)
    return PoseOnlyBatch(K=K, feat_ptr=feat_ptr, poses=poses, xyz=xyz, uv=uv)


# --------------------------------------------------------------------------- pose graphs

@dataclasses.dataclass
class PoseGraph:
    """A loop-closure pose graph (src/ssvio/loopclosing.cpp:458-532): key-frame poses T_cw, fixed flags, and
    relative-pose edges (v0, v1, measurement) with error log(M^-1 * T_v0 * T_v1^-1)."""

    poses: np.ndarray   # (N, 7) initial T_cw
    fixed: np.ndarray   # (N,) uint8
    v0: np.ndarray      # (E,) int32
    v1: np.ndarray      # (E,) int32
    meas: np.ndarray    # (E, 7)


def make_pose_graph(n_kf: int, *, seed: int = 42, n_loops: int = 3, drift_t: float = 0.03, drift_r: float = 0.004,
                    meas_sigma_t: float = 0.01, meas_sigma_r: float = 0.001, n_fixed_tail: int = 4) -> PoseGraph:
    """A trajectory of `n_kf` key-frames whose initial poses have accumulated odometry drift; edges between
    consecutive key-frames (kf -> last kf) plus `n_loops` loop edges from late to early key-frames, measured
    from the drift-free truth with small noise.  Key-frame 0, the loop key-frames' targets and the last
    `n_fixed_tail` (the "active" ones) are fixed, like loopclosing.cpp:480-486."""
    rng = _Rng(seed)
    idx = np.arange(n_kf, dtype=np.float64)
    gt = se3_exp(np.stack([0.4 * np.sin(0.05 * idx), 0.01 * idx, -1.0 * idx, 0.002 * idx, 0.01 * np.sin(0.1 * idx), 0.002 * idx], axis=1))
    poses = gt.copy()
    acc = np.zeros(6)
    for i in range(1, n_kf):   # random-walk drift
        acc = acc + np.concatenate([drift_t * rng.normal(3), drift_r * rng.normal(3)])
        poses[i] = se3_mul(se3_exp(acc[None, :]), gt[i:i + 1])[0]
    noise = lambda m: se3_exp(np.concatenate([meas_sigma_t * rng.normal((m, 3)), meas_sigma_r * rng.normal((m, 3))], axis=1))
    v0 = np.arange(1, n_kf, dtype=np.int32); v1 = v0 - 1
    lo = rng.integers(0, max(0, n_kf // 4), n_loops).astype(np.int32)
    hi = (n_kf - 1 - rng.integers(0, max(0, n_kf // 4), n_loops)).astype(np.int32)
    v0 = np.concatenate([v0, hi]); v1 = np.concatenate([v1, lo])
    rel = se3_mul(gt[v0], se3_inv(gt[v1]))        # T_v0 * T_v1^-1
    meas = se3_mul(noise(len(v0)), rel)
    fixed = np.zeros(n_kf, np.uint8)
    fixed[0] = 1; fixed[lo] = 1; fixed[n_kf - n_fixed_tail:] = 1
    return PoseGraph(poses=poses, fixed=fixed, v0=v0.astype(np.int32), v1=v1.astype(np.int32), meas=meas)
