"""g2o text import / export (ssvio_b200/g2o_io.py): record formats of the reference's vendored
g2o (types_six_dof_expmap.cpp:93-108,363-387; optimizable_graph.cpp:428,936)."""
import numpy as np
import pytest

from common import golden_case
from ssvio_b200 import g2o_io, synth


def test_round_trip_exact(tmp_path):
    g, _ = golden_case("small_fixed")
    path = str(tmp_path / "small_fixed.g2o")
    g2o_io.save_g2o(path, g, cam=0)
    h, info = g2o_io.load_g2o(path, g.K, ext=g.ext, cam=0)
    assert info is None  # identity information everywhere (backend.cpp:161)
    sel = g.cam_idx == 0
    np.testing.assert_array_equal(h.pose_idx, g.pose_idx[sel])
    np.testing.assert_array_equal(h.point_idx, g.point_idx[sel])
    np.testing.assert_array_equal(h.uv, g.uv[sel])
    np.testing.assert_array_equal(h.points, g.points)
    np.testing.assert_array_equal(h.pose_fixed, g.pose_fixed)
    np.testing.assert_array_equal(h.point_fixed, g.point_fixed)
    # poses go through two inversions (the file stores camera-to-world)
    np.testing.assert_allclose(h.poses, g.poses, rtol=0, atol=1e-14)


def test_reads_stock_records(tmp_path):
    path = tmp_path / "two.g2o"
    # camera at world (1, 2, 3), no rotation: T_cw has t = (-1, -2, -3)
    path.write_text("VERTEX_SE3:EXPMAP 7 1 2 3 0 0 0 1\nFIX 7\nVERTEX_XYZ 9 0.5 0.25 4\n"
                    "EDGE_SE3_PROJECT_XYZ:EXPMAP 9 7 600.5 180.25 2 0 2\n")
    g, info = g2o_io.load_g2o(str(path), [synth.FX, 0, synth.CX, 0, synth.FY, synth.CY, 0, 0, 1])
    assert g.n_poses == 1 and g.n_points == 1 and g.n_edges == 1
    np.testing.assert_allclose(g.poses[0], [0, 0, 0, 1, -1, -2, -3])
    assert g.pose_fixed[0] == 1 and g.point_fixed[0] == 0
    np.testing.assert_array_equal(info, [[2.0, 0.0, 2.0]])
    np.testing.assert_array_equal(g.uv, [[600.5, 180.25]])


def test_rejects_unknown_records(tmp_path):
    path = tmp_path / "bad.g2o"
    path.write_text("VERTEX_SE2 0 0 0 0\n")
    with pytest.raises(ValueError):
        g2o_io.load_g2o(str(path), np.eye(3).ravel())


@pytest.mark.gpu
def test_imported_monocular_graph_parity(tmp_path, ssba_lib, port_oracle):
    """A graph that went through the g2o text format optimises on the B200 like the oracle does
    on the same arrays (left-camera edges only: the stock g2o edge type has no camera index)."""
    from ssvio_b200 import ba
    g0, _ = golden_case("small")
    path = str(tmp_path / "small_left.g2o")
    g2o_io.save_g2o(path, g0, cam=0)
    g, _ = g2o_io.load_g2o(path, g0.K, ext=g0.ext, cam=0, iters=8)
    with ba.BundleAdjuster() as opt:
        opt.set_graph(g)
        rep = opt.optimize(g.iters)
        poses = opt.poses()
    port = port_oracle.optimize(g, jacobian="analytic")
    assert rep.iterations == port["report"].iterations
    assert abs(rep.chi2_robust - port["report"].chi2_robust) / port["report"].chi2_robust < 1e-9
    np.testing.assert_allclose(poses, port["poses"], rtol=0, atol=1e-7)
