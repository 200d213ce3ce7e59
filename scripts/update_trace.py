"""Developer aid: phase stamps of three CTAs of k_update (trace build; SSBA_LIB=ssvio_b200/lib/libssba_trace.so)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from ssvio_b200 import ba, synth
g = synth.make_config(sys.argv[1] if len(sys.argv) > 1 else "cfg3")
with ba.BundleAdjuster() as opt:
    opt.set_graph(g); opt.initialize_optimization()
    for _ in range(3):
        opt.reset_state(); opt.optimize_nowait_report(3)
    lib = ba.load_library()
    out = np.zeros(4096, dtype=np.int64)
    lib.ssba_debug_solver_trace(out.ctypes.data_as(C.POINTER(C.c_longlong)), 4096)
names = ["entered", "static part done", "predecessor complete (griddepcontrol.wait)", "W^T x per pair", "landmarks moved", "trial state linearised (+ folds)", "block sums"]
t0 = min(out[3800 + 16 * k] for k in range(3) if out[3800 + 16 * k])
for k, cta in enumerate((0, 300, 700)):
    v = out[3800 + 16 * k: 3800 + 16 * k + 7]
    print(f"CTA {cta:3d} (ns since the first of the three entered): " + " | ".join(f"{n} {int(x - t0)}" for n, x in zip(names, v)))
