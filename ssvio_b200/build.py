"""Builds ssvio_b200/lib/libssba.so (CUDA kernels + C ABI) in-tree for sm_100a with nvcc.

nvcc cross-compiles without a GPU, so this runs in the CPU container; the resulting .so travels
to the GPU box with the repository snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIB_DIR, "libssba.so")

SOURCES = ["ssba_kernels.cu", "ssba_tree_solve.cu", "ssba_pose_only.cu", "ssba_api.cu", "ssba_structure.cpp",
           "ssba_tree_program.cpp"]
HEADERS = ["ssba_geometry.cuh", "ssba_device.hpp", "ssba_structure.hpp", "ssba_solver_layout.hpp",
           "ssba_tree_program.hpp", "ssba_block_inverse.cuh"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [
        os.path.join(ROOT, "include", "ssba.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    os.makedirs(LIB_DIR, exist_ok=True)
    extra = os.environ.get("SSBA_EXTRA_NVCC_FLAGS", "").split()
    cmd = [nvcc_path()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + [
        "-I" + os.path.join(ROOT, "include"), "-I" + CSRC,
    ] + [os.path.join(CSRC, f) for f in SOURCES] + ["-o", LIB, "-ldl"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose:
        sys.stderr.write(res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
