#!/bin/sh
# Developer aid (CPU only): tests/cpp/test_structure.cpp under ThreadSanitizer with the parallel paths of
# build_structure forced on for the small test graphs (SSBA_HOST_PAR_MIN).
set -e
cd "$(dirname "$0")/../.."
g++ -O1 -g -fsanitize=thread -std=c++17 -I/usr/local/cuda/include -Iinclude -Issvio_b200/csrc tests/cpp/test_structure.cpp \
  ssvio_b200/csrc/ssba_structure.cpp ssvio_b200/csrc/ssba_tree_program.cpp -o /tmp/ssba_ts_tsan -lpthread
SSBA_HOST_PAR_MIN=64 SSBA_HOST_THREADS=6 /tmp/ssba_ts_tsan 2>&1 | grep -E "WARNING: ThreadSanitizer|SUMMARY|^OK|FAIL" | sort | uniq -c
