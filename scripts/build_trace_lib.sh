#!/bin/sh
# Developer aid: libssba_trace.so = libssba.so with the per-level clock64 trace of k_reduced_solve
# (SSBA_LIB=ssvio_b200/lib/libssba_trace.so python scripts/solver_trace.py cfg3)
cd "$(dirname "$0")/.." && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared \
  -DSSBA_SOLVER_TRACE -Iinclude -Issvio_b200/csrc ssvio_b200/csrc/ssba_kernels.cu ssvio_b200/csrc/ssba_pose_only.cu ssvio_b200/csrc/ssba_api.cu \
  ssvio_b200/csrc/ssba_structure.cpp -o ssvio_b200/lib/libssba_trace.so -ldl
