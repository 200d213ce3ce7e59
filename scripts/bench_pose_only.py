"""Developer aid: throughput of the batched pose-only LM against the compiled reference
(frames / s; the reference runs frame by frame on one host core like the front-end thread)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ssvio_b200 import ba, synth
nf = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
b = synth.make_pose_only(nf, 150, seed=1)
with ba.BundleAdjuster() as opt:
    for _ in range(3):
        out = opt.pose_only_optimize(b)
    t = time.perf_counter()
    reps = 10
    for _ in range(reps):
        out = opt.pose_only_optimize(b)
    dt = (time.perf_counter() - t) / reps
print(f"libssba: {nf} frames x ~150 features, host buffers in/out: {dt * 1e3:.2f} ms per batch = {nf / dt:,.0f} frames/s")
try:
    from oracle import bindings
    if bindings.RefOracle.available():
        ref = bindings.RefOracle()
        sub = synth.PoseOnlyBatch(K=b.K, feat_ptr=b.feat_ptr[:257], poses=b.poses[:256], xyz=b.xyz[:b.feat_ptr[256]], uv=b.uv[:b.feat_ptr[256]])
        t = time.perf_counter(); r = bindings.ref_pose_only(ref.lib, sub); dt_ref = time.perf_counter() - t
        print(f"reference (g2o, 1 core): 256 frames in {dt_ref * 1e3:.1f} ms = {256 / dt_ref:,.0f} frames/s")
        print("inlier counts equal on the first 256 frames:", bool((r[2] == out[2][:256]).all()))
except Exception as e:  # the reference .so is optional here
    print("reference not available:", e)
