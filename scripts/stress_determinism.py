"""Developer aid: repeats optimize() on a few graphs and reports any run whose trace departs from
the first one by more than rounding (rare races show up here long before they fail a test)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from ssvio_b200 import ba, synth
from common import golden_case

def rel(a, b): return abs(a - b) / max(abs(b), 1e-300)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
cases = [("tiny_long", golden_case("tiny_long")[0]), ("small", golden_case("small")[0]), ("cfg2", synth.make_config("cfg2"))]
# like the test-suite: big problems first in the same process, so that the small ones get recycled
# device memory (an uninitialised read shows up as a deviating run)
for big in ("cfg3", "cfg5"):
    with ba.BundleAdjuster() as opt:
        g = synth.make_config(big)
        opt.set_graph(g)
        opt.optimize(3)
bad = 0
for name, g in cases:
    ref = None
    for rep in range(reps):
        with ba.BundleAdjuster() as opt:
            opt.set_graph(g)
            r = opt.optimize(g.iters)
            tr = r.trace()
        if ref is None:
            ref = (r.chi2_robust, tr); continue
        n = min(len(tr), len(ref[1]), 10)
        dev = max(rel(tr[i][0], ref[1][i][0]) for i in range(n))
        if dev > 1e-9 or rel(r.chi2_robust, ref[0]) > 1e-9 or r.cholesky_failures:
            bad += 1
            print(name, "run", rep, "deviates: first-10 dev", dev, "final", r.chi2_robust, "vs", ref[0], "iters", r.iterations, "chol_fail", r.cholesky_failures)
            for i in range(min(len(tr), 6)): print("   ", tr[i], ref[1][i])
    print(name, "done", reps, "runs; final chi2", ref[0])
# one handle fed alternating graphs (arena / staging reuse and regrowth, set_graph on a live handle)
with ba.BundleAdjuster() as opt:
    refs = {}
    for rep in range(reps):
        name, g = cases[rep % len(cases)]
        opt.set_graph(g)
        r = opt.optimize(g.iters)
        if name not in refs:
            refs[name] = r.chi2_robust
        elif rel(r.chi2_robust, refs[name]) > 1e-9:
            bad += 1
            print("reused handle:", name, "run", rep, r.chi2_robust, "vs", refs[name])
    print("reused handle done", reps, "runs")
print("BAD RUNS", bad)
