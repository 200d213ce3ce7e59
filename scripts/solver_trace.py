"""Developer aid: per-level clock64 trace of k_reduced_solve (needs a -DSSBA_SOLVER_TRACE build)."""
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
    out = np.zeros(4096, dtype=np.int64)
    lib.ssba_debug_solver_trace(out.ctypes.data_as(C.POINTER(C.c_longlong)), 4096)
    info = opt.problem_info()
t0 = out[0]
nz = np.nonzero(out[:2900])[0].max()
rel = out[:nz + 1] - t0
nseg = (nz - 2) // 4
print("segments", nseg, "total cycles", rel[nz])
for sg in range(nseg):
    a, b, c = rel[1 + 3 * sg], rel[2 + 3 * sg], rel[3 + 3 * sg]
    print(f"seg {sg:3d}: start {a:8d} items {b - a:7d} wait+barrier {c - b:7d}")
base = 3 * nseg + 1
prev = rel[3 * nseg]
for k in range(1, nseg):
    v = rel[base + k]
    print(f"bwd {k:3d}: {v - prev:7d}")
    prev = v
print("tail", rel[nz] - prev)
for label, base, sgi in (("seg 6", 3000, 6), ("seg 18", 3512, 18)):
    start = out[1 + 3 * sgi]
    print(label, "rounds 0..15 of the level (cycles since level start): begin | pairs done | rows->smem | chol+inv done | published | round end")
    for rd in range(16):
        v = out[base + 16 * rd: base + 16 * rd + 7]
        if v[0] == 0: continue
        f = lambda x: int(x - start) if x else -1
        print(f"   rd {rd:2d}: {f(v[0]):6d} | {f(v[1]):6d} | {f(v[2]):6d} | {f(v[3]):6d} | {f(v[4]):6d} | {f(v[6]):6d}")

