"""Developer aid: phase / step trace of k_tree_solve (needs scripts/build_trace_lib.sh; run with
SSBA_LIB=ssvio_b200/lib/libssba_trace.so)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from ssvio_b200 import ba, synth
name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
g = synth.make_config(name)
with ba.BundleAdjuster() as opt:
    opt.set_graph(g); opt.initialize_optimization()
    for _ in range(3):
        opt.reset_state(); opt.optimize_nowait_report(2)
    lib = ba.load_library()
    out = np.zeros((16, 256), dtype=np.int64)
    lib.ssba_debug_tree_trace(out.ctypes.data_as(C.POINTER(C.c_longlong)))
    info = opt.problem_info()
print(f"{name}: cluster {info.solve_cluster} steps {info.solver_steps} top cols {info.solver_top_cols} smem {info.solver_smem_bytes}")
names = ["start", "loaded", "fwd subtrees", "barrier 1", "top contributions added", "top forward", "top backward", "x down + barrier 2", "bwd subtrees", "epilogue"]
t0 = out[:info.solve_cluster, 0].min()
for c in range(info.solve_cluster):
    g_ = out[c, :10] - t0
    print(f"CTA {c:2d} (ns since first start): " + " | ".join(f"{names[i]} {g_[i]}" for i in range(10)))
for c in (0, 1):
    cl = out[c, 16:]
    nz = np.nonzero(cl)[0]
    if len(nz) == 0: continue
    cl = cl[:nz.max() + 1]
    d = np.diff(cl)
    print(f"CTA {c} clock64 deltas (forward: interval 1, interval 2 per step; then one per backward step):")
    print("   ", [int(x) for x in d])

for c in (0, 1):
    for st in (4, 5):
        w = out[c, 128 + 32 * (st - 4): 128 + 32 * (st - 4) + 30].reshape(6, 5)
        if w[0, 0] == 0: continue
        t0 = w[:, 0].min()
        print(f"CTA {c} step {st} (cycles since the first warp entered the step); warp 0 = diagonal: start | own blocks scaled | products done | inverse done | barrier passed;  warps 1..5 = look-ahead: start | panel done | named barrier passed | look-ahead done | barrier passed")
        for k in range(6):
            print(f"    warp {k}: " + " | ".join(str(int(x - t0)) if x else "-" for x in w[k]))
