"""Developer aid: where the end-to-end time of one set_graph + optimize + read-back goes."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ssvio_b200 import ba, synth
g = synth.make_config(sys.argv[1] if len(sys.argv) > 1 else "cfg3")
with ba.BundleAdjuster() as opt:
    for rep in range(4):
        t0 = time.perf_counter(); opt.set_graph(g)
        t1 = time.perf_counter(); opt.initialize_optimization()
        t2 = time.perf_counter(); r = opt.optimize(g.iters)
        t3 = time.perf_counter(); p = opt.poses(); q = opt.points()
        t4 = time.perf_counter()
        print(f"set_graph {1e3*(t1-t0):.3f}  initialize {1e3*(t2-t1):.3f}  optimize {1e3*(t3-t2):.3f}  read-back {1e3*(t4-t3):.3f}  total {1e3*(t4-t0):.3f} ms")
