"""The N > 1 path on CPU (gloo, world_size 2): the landmark shard plan of the product covers every
landmark exactly once and balances the edges, and the per-shard reduced pose systems (computed
here by the oracle) all-reduce to the full system — the identity the NCCL path relies on."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _reduced_system(lib, g, lam, owner, rank):
    nfp = C.c_int32(0)
    nmax = g.n_poses
    S = np.zeros((6 * nmax) ** 2)
    b = np.zeros(6 * nmax)
    dp, ip, bp = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_uint8)
    f = lib.ssba_oracle_reduced_system
    f.restype = C.c_int
    rc = f(g.K.ctypes.data_as(dp), C.c_int32(g.ext.shape[0]), g.ext.ctypes.data_as(dp),
           C.c_int32(g.n_poses), g.poses.ctypes.data_as(dp), g.pose_fixed.ctypes.data_as(bp),
           C.c_int32(g.n_points), g.points.ctypes.data_as(dp), g.point_fixed.ctypes.data_as(bp),
           C.c_int32(g.n_edges), g.pose_idx.ctypes.data_as(ip), g.point_idx.ctypes.data_as(ip),
           g.cam_idx.ctypes.data_as(bp), g.uv.ctypes.data_as(dp), C.c_double(g.huber_delta),
           C.c_double(lam), owner.ctypes.data_as(ip) if owner is not None else None, C.c_int32(rank),
           C.c_int32(nmax), S.ctypes.data_as(dp), b.ctypes.data_as(dp), C.byref(nfp))
    assert rc == 0
    n = 6 * nfp.value
    return S[: n * n].reshape(n, n), b[:n]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    from oracle import bindings
    from ssvio_b200 import ba, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lib = bindings.PortOracle().lib
        g = synth.make_config("small", seed=9, fix_first_pose=True, n_fixed_points=12)
        owner = ba.plan_shards(g, world)   # host-only entry of the product library
        # rank 0 broadcasts its plan: every rank must have computed the same one
        t = torch.from_numpy(owner.copy())
        dist.broadcast(t, 0)
        assert np.array_equal(t.numpy(), owner)
        lam = 37.5
        S_r, b_r = _reduced_system(lib, g, lam, owner, rank)
        ts, tb = torch.from_numpy(S_r.copy()), torch.from_numpy(b_r.copy())
        dist.all_reduce(ts)
        dist.all_reduce(tb)
        S_full, b_full = _reduced_system(lib, g, lam, None, 0)
        np.testing.assert_allclose(ts.numpy(), S_full, rtol=1e-12, atol=1e-9 * np.abs(S_full).max())
        np.testing.assert_allclose(tb.numpy(), b_full, rtol=1e-12, atol=1e-9 * np.abs(b_full).max())
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_shard_plan_properties(ssba_lib):
    from ssvio_b200 import ba, synth
    g = synth.make_config("cfg1", fix_first_pose=True, n_fixed_points=30)
    active = np.zeros(g.n_points, bool)
    both_fixed = g.pose_fixed[g.pose_idx].astype(bool) & g.point_fixed[g.point_idx].astype(bool)
    active[g.point_idx[~both_fixed]] = True
    for world in (1, 2, 4, 8):
        owner = ba.plan_shards(g, world)
        assert np.array_equal(owner >= 0, active)          # every active landmark has one owner
        assert owner.max() == world - 1
        edges = np.bincount(owner[g.point_idx[~both_fixed]], minlength=world)
        assert edges.sum() == (~both_fixed).sum()
        assert edges.max() - edges.min() <= 0.1 * edges.mean() + 20   # balanced by edge count
    assert np.array_equal(ba.plan_shards(g, 1)[active], np.zeros(active.sum(), np.int32))


def test_reduced_system_is_additive_over_shards_gloo(port_oracle, ssba_lib):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, 29533, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
