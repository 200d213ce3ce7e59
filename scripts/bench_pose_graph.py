"""Developer aid: the loop closer's pose-graph optimisation (optimize(20)) on libssba against the compiled
reference on one host core, from host buffers."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ssvio_b200 import ba, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 400
pg = synth.make_pose_graph(n, seed=9, n_loops=8)
with ba.BundleAdjuster() as opt:
    for _ in range(3):
        poses, rep = opt.pose_graph_optimize(pg)
    t = time.perf_counter(); poses, rep = opt.pose_graph_optimize(pg); dt = time.perf_counter() - t
print(f"libssba: {n} key-frames, {len(pg.v0)} edges, optimize(20): {dt * 1e3:.2f} ms, chi2 {rep.chi2_initial:.6g} -> {rep.chi2_robust:.6g}, iterations {rep.iterations}")
try:
    from oracle import bindings
    if bindings.RefOracle.available():
        ref = bindings.RefOracle()
        t = time.perf_counter(); P, r = bindings.ref_pose_graph(ref.lib, pg); dt = time.perf_counter() - t
        print(f"reference (g2o + LinearSolverEigen, 1 core): {dt * 1e3:.2f} ms, chi2 -> {r.chi2_robust:.6g}, iterations {r.iterations}")
except Exception as e:
    print("reference not available:", e)
