"""Developer aid: one optimize() of a synthetic config through the C ABI (for compute-sanitizer / ncu runs).
usage: python scripts/run_case.py <config> [iters]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ssvio_b200 import ba, synth  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "small"
g = synth.make_config(name)
iters = int(sys.argv[2]) if len(sys.argv) > 2 else g.iters
with ba.BundleAdjuster() as opt:
    opt.set_graph(g)
    rep = opt.optimize(iters)
    info = opt.problem_info()
    print(f"{name}: iterations {rep.iterations} chi2 {rep.chi2_robust:.9f} failures {rep.cholesky_failures} "
          f"solver_kind {info.solver_kind} cluster {info.solve_cluster} steps {info.solver_steps} smem {info.solver_smem_bytes}")
    print("trace", [(round(c, 6), l, t) for c, l, t in rep.trace()])
