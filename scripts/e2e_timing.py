"""Developer aid: where the end-to-end time of one set_graph + optimize + read-back goes
(median over repetitions).  SSBA_TIMING=1 adds the library's own section timers on stderr."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ssvio_b200 import ba, synth
g = synth.make_config(sys.argv[1] if len(sys.argv) > 1 else "cfg3")
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
rows = []
with ba.BundleAdjuster() as opt:
    for rep in range(reps):
        if os.environ.get('E2E_FRESH', '1') != '0': opt.drop_structure()
        t0 = time.perf_counter(); opt.set_graph(g)
        t1 = time.perf_counter(); opt.initialize_optimization()
        t2 = time.perf_counter(); r = opt.optimize(g.iters)
        t3 = time.perf_counter(); p = opt.poses(); q = opt.points()
        t4 = time.perf_counter()
        rows.append([t1 - t0, t2 - t1, t3 - t2, t4 - t3, t4 - t0])
m = 1e3 * np.median(np.array(rows[3:]), axis=0)
print(f"median of {reps - 3}: set_graph {m[0]:.3f}  initialize {m[1]:.3f}  optimize {m[2]:.3f}  read-back {m[3]:.3f}  total {m[4]:.3f} ms")
