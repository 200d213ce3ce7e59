"""Developer aid: run one config through libssba and print it next to the golden/oracle values."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from ssvio_b200 import ba, synth  # noqa: E402
from common import golden_scalars  # noqa: E402


def main():
    names = sys.argv[1:] or ["tiny"]
    gold = golden_scalars()
    for name in names:
        g = synth.make_config(name)
        with ba.BundleAdjuster(profile=True) as opt:
            t0 = time.perf_counter()
            opt.set_graph(g)
            opt.initialize_optimization()
            t1 = time.perf_counter()
            rep = opt.optimize(g.iters)
            t2 = time.perf_counter()
            info = opt.problem_info()
            print(f"== {name}: fp={info.n_free_poses} fl={info.n_free_points} edges={info.n_active_edges} "
                  f"pairs={info.n_pairs} schur={info.n_schur_blocks} factor={info.n_factor_blocks} "
                  f"bytes={info.device_bytes}")
            print(f"   its={rep.iterations} chi0={rep.chi2_initial:.6f} chi={rep.chi2_robust:.10f} plain={rep.chi2_plain:.6f} "
                  f"lam={rep.lambda_:.6g} cholfail={rep.cholesky_failures} setup={1e3*(t1-t0):.2f}ms opt={1e3*(t2-t1):.2f}ms")
            for t in rep.trace():
                print("     %.8f %.6g %d" % t)
            if name in gold:
                ga = gold[name]["analytic"]
                print(f"   golden analytic chi={ga['chi2_robust']:.10f} rel={abs(rep.chi2_robust-ga['chi2_robust'])/ga['chi2_robust']:.2e}"
                      f"  numeric rel={abs(rep.chi2_robust-gold[name]['numeric']['chi2_robust'])/ga['chi2_robust']:.2e}")
            p = opt.profile()
            print(f"   profile ms: lin={p.ms_linearize:.3f}/{p.n_linearize} schur={p.ms_schur:.3f}/{p.n_schur} "
                  f"solve={p.ms_reduced_solve:.3f}/{p.n_reduced_solve} update={p.ms_update_chi2:.3f}/{p.n_update_chi2} launches={p.kernel_launches}")
            # steady-state timing
            for _ in range(3):
                opt.reset_state(); opt.optimize_nowait_report(g.iters)
            t3 = time.perf_counter()
            n = 10
            for _ in range(n):
                opt.reset_state(); opt.optimize_nowait_report(g.iters)
            t4 = time.perf_counter()
            print(f"   resident optimize({g.iters}): {1e3*(t4-t3)/n:.3f} ms -> {g.iters*n/(t4-t3):.1f} LM it/s")


if __name__ == "__main__":
    main()
