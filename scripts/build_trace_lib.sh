#!/bin/sh
# Developer aid: libssba_trace.so = libssba.so with the clock traces of the reduced solvers
# (SSBA_LIB=ssvio_b200/lib/libssba_trace.so python scripts/solver_trace.py cfg3   -- k_reduced_solve, SSBA_SOLVER=level
#  SSBA_LIB=ssvio_b200/lib/libssba_trace.so python scripts/tree_trace.py cfg3     -- k_tree_solve)
cd "$(dirname "$0")/.." && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared \
  -DSSBA_SOLVER_TRACE -Iinclude -Issvio_b200/csrc ssvio_b200/csrc/ssba_kernels.cu ssvio_b200/csrc/ssba_tree_solve.cu ssvio_b200/csrc/ssba_pose_only.cu ssvio_b200/csrc/ssba_api.cu \
  ssvio_b200/csrc/ssba_structure.cpp ssvio_b200/csrc/ssba_tree_program.cpp -o ssvio_b200/lib/libssba_trace.so -ldl
