"""Parity of the CUDA path (through the C ABI) against the oracle — run on the B200 with -m gpu.

Checkers: golden fixtures made by the compiled reference (tests/golden), the C restatement
(oracle port) on the same seeded inputs, and the compiled reference itself when the prebuilt
oracle/_ref travelled with the snapshot.  Tolerance: north_star's 1e-6 relative on the final
chi2 against the reference as shipped (numeric Jacobians); 1e-9 against the same analytic
Jacobians.
"""
import numpy as np
import pytest

from common import CHI2_RTOL, CHI2_RTOL_SAME_JACOBIAN, converging_prefix, golden_case, golden_scalars, rel
from ssvio_b200 import synth

pytestmark = pytest.mark.gpu

SMALL = ["tiny", "small", "small_fixed", "cfg1"]


@pytest.fixture(scope="module")
def BA(ssba_lib):
    from ssvio_b200 import ba
    return ba.BundleAdjuster


def run_gpu(BA, g, iters=None, **kw):
    with BA(**kw) as opt:
        opt.set_graph(g)
        opt.initialize_optimization()
        rep = opt.optimize(g.iters if iters is None else iters)
        return dict(report=rep, poses=opt.poses(), points=opt.points(), errors=opt.edge_errors(),
                    chi2=opt.chi2(), outliers=opt.count_outliers(5.891), info=opt.problem_info())


@pytest.mark.parametrize("name", SMALL)
def test_golden_small_cases(BA, name):
    g, z = golden_case(name)
    gold = golden_scalars()[name]
    r = run_gpu(BA, g)
    rep = r["report"]
    assert rep.iterations == gold["analytic"]["iterations"]
    assert rel(rep.chi2_initial, gold["analytic"]["chi2_initial"]) < 1e-12
    # vs the reference as shipped (numeric Jacobians): the north_star bar
    assert rel(rep.chi2_robust, gold["numeric"]["chi2_robust"]) < CHI2_RTOL
    # vs the reference with the same analytic Jacobians: tight
    assert rel(rep.chi2_robust, gold["analytic"]["chi2_robust"]) < CHI2_RTOL_SAME_JACOBIAN
    assert rel(rep.chi2_plain, gold["analytic"]["chi2_plain"]) < CHI2_RTOL_SAME_JACOBIAN
    for (chi, lam, trials), (gchi, glam, gtrials) in zip(rep.trace(), gold["analytic"]["trace"]):
        assert rel(chi, gchi) < CHI2_RTOL_SAME_JACOBIAN and rel(lam, glam) < 1e-6 and trials == gtrials
    np.testing.assert_allclose(r["poses"], z["analytic_poses"], rtol=0, atol=1e-7)
    np.testing.assert_allclose(r["points"], z["analytic_points"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(r["errors"], z["analytic_errors"], rtol=0, atol=1e-5)
    # read-outs are consistent with the report
    assert rel(r["chi2"][0], rep.chi2_plain) < 1e-12 and rel(r["chi2"][1], rep.chi2_robust) < 1e-12
    e2 = (r["errors"] ** 2).sum(1)
    assert r["outliers"] == (int((e2 > 5.891).sum()), int((e2 <= 5.891).sum()))


@pytest.mark.parametrize("name", ["tiny", "small_fixed"])
def test_numeric_jacobian_mode(BA, name):
    """jacobian_mode = NUMERIC reproduces the reference as shipped (central differences)."""
    g, _ = golden_case(name)
    gold = golden_scalars()[name]["numeric"]
    rep = run_gpu(BA, g, jacobian="numeric")["report"]
    assert rep.iterations == gold["iterations"]
    assert rel(rep.chi2_robust, gold["chi2_robust"]) < CHI2_RTOL


@pytest.mark.parametrize("name", ["cfg2", "cfg3"])
def test_baseline_configs_against_golden(BA, name):
    g = synth.make_config(name)
    gold = golden_scalars()[name]
    assert g.n_edges == gold["n_edges"]
    rep = run_gpu(BA, g)["report"]
    assert rep.iterations == gold["numeric"]["iterations"]
    assert rel(rep.chi2_robust, gold["numeric"]["chi2_robust"]) < CHI2_RTOL
    assert rel(rep.chi2_robust, gold["analytic"]["chi2_robust"]) < CHI2_RTOL_SAME_JACOBIAN
    for (chi, lam, trials), (gchi, glam, gtrials) in zip(rep.trace(), gold["analytic"]["trace"]):
        assert rel(chi, gchi) < 1e-8 and trials == gtrials


def test_cfg5_global_ba_fixed_gauge(BA):
    """500 KF / 100k landmarks / 1M edges, KF0 fixed (BASELINE config 5) on one GPU."""
    g = synth.make_config("cfg5")
    gold = golden_scalars()["cfg5"]
    rep = run_gpu(BA, g)["report"]
    assert rep.iterations == 10
    assert rel(rep.chi2_robust, gold["numeric"]["chi2_robust"]) < CHI2_RTOL
    assert rel(rep.chi2_robust, gold["analytic"]["chi2_robust"]) < 1e-8


@pytest.mark.parametrize("seed", [3, 11])
def test_against_port_oracle_fresh_seeds(BA, port_oracle, seed):
    g = synth.make_config("small", seed=seed, fix_first_pose=bool(seed % 2), n_fixed_points=seed)
    a = port_oracle.optimize(g, jacobian="analytic")
    r = run_gpu(BA, g)
    assert r["report"].iterations == a["report"].iterations
    assert rel(r["report"].chi2_robust, a["report"].chi2_robust) < CHI2_RTOL_SAME_JACOBIAN
    np.testing.assert_allclose(r["poses"], a["poses"], atol=1e-7)
    np.testing.assert_allclose(r["points"], a["points"], atol=1e-6)


def test_against_compiled_reference_live(BA, ref_oracle):
    g = synth.make_config("cfg1", seed=5)
    a = ref_oracle.optimize(g, jacobian="numeric")  # as shipped
    r = run_gpu(BA, g)
    assert r["report"].iterations == a["report"].iterations
    assert rel(r["report"].chi2_robust, a["report"].chi2_robust) < CHI2_RTOL


def test_long_run_rejections_and_terminate(BA):
    """lambda decays until steps get rejected and LM terminates (levenberg.cpp:137-148).

    The graph has no fixed vertex, so once lambda is ~1e-7 the reduced system is singular up to
    rounding and the accept / reject decisions of the last few iterations are rounding noise
    (in the reference too).  The trajectory is therefore compared strictly over the converging
    prefix, and the chaotic tail only through what must hold anyway: it terminates early after
    rejections, and ends at the reference's minimum within the north-star tolerance."""
    g, _ = golden_case("tiny_long")
    gold = golden_scalars()["tiny_long"]["analytic"]
    n = converging_prefix(gold["trace"])
    assert n >= 10
    # strict: the first n iterations, run on their own
    rep = run_gpu(BA, g, iters=n)["report"]
    tr = rep.trace()
    assert rep.iterations == n, tr
    for i, ((chi, lam, trials), (gchi, glam, gtrials)) in enumerate(zip(tr, gold["trace"][:n])):
        assert rel(chi, gchi) < 1e-9 and rel(lam, glam) < 1e-6 and trials == gtrials, f"iteration {i}: {tr}"
    assert rel(rep.chi2_robust, gold["trace"][n - 1][0]) < CHI2_RTOL_SAME_JACOBIAN
    # the full run: same prefix, then rejections and an early Terminate
    rep = run_gpu(BA, g)["report"]
    tr = rep.trace()
    ctx = f"iterations={rep.iterations} chol_fail={rep.cholesky_failures} chi2={rep.chi2_robust!r} trace={tr}"
    for i, ((chi, lam, trials), (gchi, glam, gtrials)) in enumerate(list(zip(tr, gold["trace"]))[:n]):
        assert rel(chi, gchi) < 1e-9 and trials == gtrials, f"iteration {i}: {ctx}"
    assert rep.iterations < g.iters and rep.last_result == 2, ctx   # SSBA_SOLVER_TERMINATE
    # Terminate is only reached through the reject branch (10 failed trials, rho == 0 exactly, or
    # lambda overflow, levenberg.cpp:137-148): the last trial raised lambda
    assert tr[-1][2] > 1 or tr[-1][1] > tr[-2][1], ctx
    assert rel(rep.chi2_robust, gold["chi2_robust"]) < CHI2_RTOL, ctx
    # ... and the chaotic tail is chaotic in the arithmetic only, not from run to run: the same decisions again
    rep2 = run_gpu(BA, g)["report"]
    assert rep2.trace() == tr and rep2.iterations == rep.iterations and rep2.chi2_robust == rep.chi2_robust, (ctx, rep2.trace())


def test_rejected_trials_match_oracle(BA, port_oracle):
    """Decisive rejections: with tau = 1e-12 LM starts as Gauss-Newton on a gauge-fixed graph with
    badly perturbed landmarks; a few iterations in, trials overshoot and are rejected several
    times in a row (re-solve with a larger lambda, no re-linearisation; levenberg.cpp:119-145).
    The trajectory, trial counts included, must match the oracle's."""
    g, _ = golden_case("small_fixed")
    rng = np.random.default_rng(5)
    g.points = g.points + 1.5 * rng.standard_normal(g.points.shape) * (1 - g.point_fixed[:, None])
    g.iters = 10
    port_oracle.set_tau(1e-12)
    try:
        port = port_oracle.optimize(g, jacobian="analytic")["report"]
    finally:
        port_oracle.set_tau(1e-5)
    ptr = port.trace()
    assert max(t[2] for t in ptr) > 1
    rep = run_gpu(BA, g, tau=1e-12)["report"]
    tr = rep.trace()
    assert rep.iterations == port.iterations and len(tr) == len(ptr), (tr, ptr)
    for i, ((chi, lam, trials), (pchi, plam, ptrials)) in enumerate(zip(tr, ptr)):
        assert rel(chi, pchi) < 1e-8 and rel(lam, plam) < 1e-6 and trials == ptrials, f"iteration {i}: {tr} vs {ptr}"
    assert rel(rep.chi2_robust, port.chi2_robust) < 1e-8


def dense_graph(n_kf=140, n_points=6, seed=12, iters=4):
    """Key-frames 5 cm apart that ALL see the same few landmarks in both cameras (a slow approach to a wall)."""
    rng = np.random.default_rng(seed)
    idx = np.arange(n_kf, dtype=np.float64)
    gt = synth.se3_exp(np.stack([0.001 * idx, 0 * idx, -0.05 * idx, 0 * idx, 0.0005 * idx, 0 * idx], axis=1))
    pts = np.stack([rng.uniform(-6, 6, n_points), rng.uniform(-2, 2, n_points), rng.uniform(14, 24, n_points)], axis=1)
    K = np.array([synth.FX, 0, synth.CX, 0, synth.FY, synth.CY, 0, 0, 1.0])
    ext = np.array([[0, 0, 0, 1, 0, 0, 0], [0, 0, 0, 1, -synth.BF / synth.FX, 0, 0]], dtype=np.float64)
    pose_idx, point_idx, cam_idx, uv = [], [], [], []
    for j in range(n_points):
        for i in range(n_kf):
            pb = synth.se3_act(gt[i], pts[j])
            for c in range(2):
                pc = synth.se3_act(ext[c], pb)
                pose_idx.append(i); point_idx.append(j); cam_idx.append(c)
                uv.append([synth.FX * pc[0] / pc[2] + synth.CX + 0.5 * rng.standard_normal(),
                           synth.FY * pc[1] / pc[2] + synth.CY + 0.5 * rng.standard_normal()])
    d = np.concatenate([0.02 * rng.standard_normal((n_kf, 3)), 0.002 * rng.standard_normal((n_kf, 3))], axis=1)
    poses = synth.se3_mul(synth.se3_exp(d), gt)
    return synth.Graph(K=K, ext=ext, poses=poses, pose_fixed=np.zeros(n_kf, np.uint8),
                       points=pts + 0.1 * rng.standard_normal(pts.shape), point_fixed=np.zeros(n_points, np.uint8),
                       pose_idx=np.array(pose_idx, np.int32), point_idx=np.array(point_idx, np.int32),
                       cam_idx=np.array(cam_idx, np.uint8), uv=np.array(uv), name="dense", iters=iters)


def test_landmarks_seen_by_more_poses_than_a_cta_has_threads(BA, port_oracle):
    """140 key-frames that all see the same six landmarks: 140 (pose, landmark) pairs per landmark (the strided
    big-chunk path of k_linearize / k_update), k = 140 poses per Schur run (unstaged W, 9870 block pairs in
    chunks of 32), a dense 140-column reduced system (natural order, one column per level)."""
    g = dense_graph()
    assert g.n_edges == 140 * 6 * 2
    port = port_oracle.optimize(g, jacobian="analytic")
    r = run_gpu(BA, g)
    rep = r["report"]
    assert r["info"].n_pairs == 140 * 6 and r["info"].n_schur_blocks == 140 * 141 // 2
    assert rep.iterations == port["report"].iterations
    assert rel(rep.chi2_initial, port["report"].chi2_initial) < 1e-12
    assert rel(rep.chi2_robust, port["report"].chi2_robust) < 1e-8
    np.testing.assert_allclose(r["errors"], port["errors"], rtol=0, atol=1e-5)


def test_step_api_equals_optimize(BA):
    """ssba_step(i) (what the g2o shim's solve(i) calls) walks the same trajectory as optimize(N)."""
    g, _ = golden_case("small")
    with BA() as a, BA() as b:
        a.set_graph(g); b.set_graph(g)
        rep = a.optimize(6)
        for i in range(6):
            res, rec = b.step(i)
            assert res == 1
            assert rel(rec.chi2, rep.iters[i].chi2) < 1e-13 and rel(rec.lambda_, rep.iters[i].lambda_) < 1e-13
        np.testing.assert_allclose(a.poses(), b.poses(), rtol=0, atol=1e-10)
        np.testing.assert_allclose(a.points(), b.points(), rtol=0, atol=1e-10)


def test_reset_state_and_determinism(BA):
    g = synth.make_config("cfg1")
    with BA() as opt:
        opt.set_graph(g)
        r1 = opt.optimize(5); p1 = opt.poses()
        opt.reset_state()
        np.testing.assert_array_equal(opt.poses(), g.poses)
        r2 = opt.optimize(5); p2 = opt.poses()
        # no fp64 atomics anywhere on the path (k_schur_reduce adds in a host-planned order): bit for bit
        assert r2.chi2_robust == r1.chi2_robust and r2.trace() == r1.trace()
        np.testing.assert_array_equal(p1, p2)


def test_bitwise_reproducible_across_handles(BA):
    """Two handles, the window of the headline benchmark: the same bits (the reference is deterministic too)."""
    g = synth.make_config("cfg3")
    a, b = run_gpu(BA, g), run_gpu(BA, g)
    assert a["report"].trace() == b["report"].trace()
    np.testing.assert_array_equal(a["poses"], b["poses"])
    np.testing.assert_array_equal(a["points"], b["points"])
    np.testing.assert_array_equal(a["errors"], b["errors"])


def test_edge_order_invariance(BA):
    """The reference's edge order is an unordered_map walk (backend.cpp:113,136); the result may
    only move by summation-order rounding."""
    g = synth.make_config("small")
    perm = np.random.default_rng(0).permutation(g.n_edges)
    h = synth.Graph(K=g.K, ext=g.ext, poses=g.poses, pose_fixed=g.pose_fixed, points=g.points,
                    point_fixed=g.point_fixed, pose_idx=g.pose_idx[perm].copy(), point_idx=g.point_idx[perm].copy(),
                    cam_idx=g.cam_idx[perm].copy(), uv=g.uv[perm].copy(), huber_delta=g.huber_delta, iters=g.iters)
    a, b = run_gpu(BA, g), run_gpu(BA, h)
    assert rel(a["report"].chi2_robust, b["report"].chi2_robust) < 1e-11
    np.testing.assert_allclose(a["errors"][perm], b["errors"], atol=1e-8)


def test_empty_and_degenerate(BA):
    from ssvio_b200 import ba
    g = synth.make_config("tiny")
    # nothing to optimise -> optimize() returns -1 (sparse_optimizer.cpp:368-371)
    g.pose_fixed[:] = 1; g.point_fixed[:] = 1
    with BA() as opt:
        opt.set_graph(g)
        assert opt.optimize(5).iterations == -1
        np.testing.assert_array_equal(opt.poses(), g.poses)
    # empty graph
    with BA() as opt:
        opt.set_cameras(g.K, g.ext)
        opt.set_poses(g.poses[:0]); opt.set_points(g.points[:0])
        opt.set_edges(g.pose_idx[:0], g.point_idx[:0], g.cam_idx[:0], g.uv[:0])
        assert opt.optimize(5).iterations == -1
    # bad index is an error, not a crash
    with BA() as opt:
        opt.set_graph(synth.make_config("tiny"))
        bad = synth.make_config("tiny")
        bad.pose_idx[0] = 99
        opt.set_edges(bad.pose_idx, bad.point_idx, bad.cam_idx, bad.uv)
        with pytest.raises(ba.SsbaError):
            opt.initialize_optimization()
    # optimize(0) does nothing
    with BA() as opt:
        g2 = synth.make_config("tiny")
        opt.set_graph(g2)
        assert opt.optimize(0).iterations == 0
        np.testing.assert_array_equal(opt.poses(), g2.poses)


def test_no_huber_and_information_matrix(BA, port_oracle):
    g = synth.make_config("small")
    g.huber_delta = 0.0
    a = port_oracle.optimize(g, jacobian="analytic")["report"]
    r = run_gpu(BA, g)["report"]
    assert rel(r.chi2_robust, a.chi2_robust) < CHI2_RTOL_SAME_JACOBIAN
    assert rel(r.chi2_robust, r.chi2_plain) < 1e-14  # no kernel: robust == plain
    # information = s*I scales chi2 by s and leaves the minimiser where it is (Huber off)
    with BA() as opt:
        opt.set_cameras(g.K, g.ext); opt.set_poses(g.poses, g.pose_fixed); opt.set_points(g.points, g.point_fixed)
        info = np.tile(np.array([4.0, 0.0, 4.0]), (g.n_edges, 1))
        opt.set_edges(g.pose_idx, g.point_idx, g.cam_idx, g.uv, info=info, huber_delta_all=0.0)
        rep = opt.optimize(g.iters)
    assert rel(rep.chi2_initial, 4.0 * r.chi2_initial) < 1e-12


def test_full_size_properties_cfg3(BA):
    """Size-independent properties at BASELINE's full size: chi2 decreases monotonically over
    accepted steps, the report agrees with an independent read-out, per-edge errors sum to chi2,
    fixed vertices do not move."""
    g = synth.make_config("cfg3", n_fixed_points=200, fix_first_pose=True)
    r = run_gpu(BA, g)
    rep = r["report"]
    chis = [rep.chi2_initial] + [t[0] for t in rep.trace()]
    assert all(b <= a * (1 + 1e-12) for a, b in zip(chis, chis[1:]))
    e2 = (r["errors"] ** 2).sum(1)
    assert rel(e2.sum(), rep.chi2_plain) < 1e-11
    d = g.huber_delta
    rho = np.where(e2 <= d * d, e2, 2 * np.sqrt(e2) * d - d * d)
    assert rel(rho.sum(), rep.chi2_robust) < 1e-11
    np.testing.assert_array_equal(r["poses"][0], g.poses[0])
    fixed = g.point_fixed.astype(bool)
    np.testing.assert_array_equal(r["points"][fixed], g.points[fixed])
    assert np.abs(np.linalg.norm(r["poses"][:, :4], axis=1) - 1).max() < 1e-12
    both_fixed = g.pose_fixed[g.pose_idx].astype(bool) & g.point_fixed[g.point_idx].astype(bool)
    assert r["info"].n_active_edges == g.n_edges - int(both_fixed.sum())
    np.testing.assert_array_equal(r["errors"][both_fixed], 0.0)


def test_g2o_shim_drop_in(BA):
    """The boundary: ssvio's graph construction + g2o::SparseOptimizer with
    ssba::OptimizationAlgorithmLevenbergCuda installed by setAlgorithm() (backend.cpp:83-86) gives
    the reference's result, read back through the plain g2o API (v->estimate(), e->chi2())."""
    from oracle import bindings
    if not bindings.ShimHarness.available():
        pytest.skip("oracle/_ref/libssba_shim_test.so not built (needs /root/reference at build time)")
    shim = bindings.ShimHarness()
    for name in ("small_fixed", "cfg1"):
        g, z = golden_case(name)
        gold = golden_scalars()[name]
        r = shim.optimize(g)
        rep = r["report"]
        assert rep.iterations == gold["numeric"]["iterations"]
        assert rel(rep.chi2_robust, gold["numeric"]["chi2_robust"]) < CHI2_RTOL
        assert rel(rep.chi2_robust, gold["analytic"]["chi2_robust"]) < CHI2_RTOL_SAME_JACOBIAN
        np.testing.assert_allclose(r["poses"], z["analytic_poses"], atol=1e-7)
        np.testing.assert_allclose(r["points"], z["analytic_points"], atol=1e-6)
        # e->chi2() of every ACTIVE edge, as backend.cpp:184 reads it (an edge between two fixed
        # vertices is never active; g2o leaves its _error uninitialised, in the reference too)
        act = ~(g.pose_fixed[g.pose_idx].astype(bool) & g.point_fixed[g.point_idx].astype(bool))
        np.testing.assert_allclose(r["edge_chi2"][act], (z["analytic_errors"] ** 2).sum(1)[act], rtol=1e-6, atol=1e-6)
        assert rel(r["edge_chi2"][act].sum(), rep.chi2_plain) < 1e-12


def test_outlier_rounds_like_backend(BA, ref_oracle):
    """The <= 5 rounds of initializeOptimization(); optimize(10) with the inlier-ratio rule
    (backend.cpp:175-203): many outliers keep the ratio under 0.7 so every round runs; few
    outliers stop after the first."""
    for frac, want_rounds in ((0.4, 5), (0.02, 1)):
        g = synth.make_config("small", outlier_frac=frac, seed=3)
        a = ref_oracle.optimize(g, iters=10, jacobian="numeric", trace=False, rounds=5, outlier_threshold=5.891)
        with BA() as opt:
            opt.set_graph(g)
            rounds, n_out, n_in, rep = opt.optimize_rounds(5, 10, 5.891, 0.7)
            poses = opt.poses()
        assert rounds == a["rounds"] == want_rounds
        assert n_out + n_in == g.n_edges
        assert abs(n_out - a["outliers"]) <= max(2, 0.002 * g.n_edges)   # edges sitting on the threshold may flip
        assert rel(rep.chi2_robust, a["report"].chi2_robust) < 5e-6
        # no pose is fixed (backend.cpp:93-103): after 50 iterations lambda is tiny and the
        # estimates may drift along the gauge freedom, so only chi2 is comparable here
        assert np.isfinite(poses).all()
