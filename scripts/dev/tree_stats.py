"""Developer aid (CPU only): dump a synthetic config and print the structure / k_tree_solve program statistics.
usage: python scripts/dev/tree_stats.py cfg3 [env SSBA_TREE_CLUSTER_CAP=16 ...]"""
import os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from ssvio_b200 import synth
name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
g = synth.make_config(name)
path = f"/tmp/ssba_graph_{name}.bin"
with open(path, "wb") as f:
    np.array([g.n_poses, g.n_points, g.n_edges], dtype=np.int32).tofile(f)
    g.pose_fixed.astype(np.uint8).tofile(f); g.point_fixed.astype(np.uint8).tofile(f)
    g.pose_idx.astype(np.int32).tofile(f); g.point_idx.astype(np.int32).tofile(f); g.cam_idx.astype(np.uint8).tofile(f)
exe = "/tmp/ssba_tree_stats"
csrc = os.path.join(ROOT, "ssvio_b200", "csrc")
subprocess.run(["g++", "-O2", "-std=c++17", "-I/usr/local/cuda/include", "-I" + os.path.join(ROOT, "include"), "-I" + csrc,
                os.path.join(ROOT, "scripts", "dev", "tree_stats.cpp"), os.path.join(csrc, "ssba_structure.cpp"),
                os.path.join(csrc, "ssba_tree_program.cpp"), "-o", exe, "-lpthread"], check=True)
subprocess.run([exe, path], check=True)
