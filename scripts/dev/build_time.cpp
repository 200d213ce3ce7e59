// Developer aid (CPU only): wall time of build_structure on a graph dumped by scripts/dev/tree_stats.py
// (median over repetitions; SSBA_TIMING=1 prints the sections).  usage: build_time <graph.bin> [reps]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <vector>
#include "ssba_structure.hpp"
using namespace ssba;
int main(int argc, char **argv) {
  FILE *f = std::fopen(argv[1], "rb");
  int hdr[3];
  if (!f || std::fread(hdr, 4, 3, f) != 3) return 1;
  const int reps = argc > 2 ? std::atoi(argv[2]) : 30;
  HostGraph g;
  g.have_cams = true; g.cams.n = 2;
  g.n_poses = hdr[0]; g.n_points = hdr[1]; g.n_edges = hdr[2];
  g.pose_fixed.resize(g.n_poses); g.point_fixed.resize(g.n_points);
  g.e_pose.resize(g.n_edges); g.e_point.resize(g.n_edges); g.e_cam.resize(g.n_edges);
  g.values_on_device = true;
  size_t ok = std::fread(g.pose_fixed.data(), 1, g.n_poses, f) + std::fread(g.point_fixed.data(), 1, g.n_points, f) +
              std::fread(g.e_pose.data(), 4, g.n_edges, f) + std::fread(g.e_point.data(), 4, g.n_edges, f) + std::fread(g.e_cam.data(), 1, g.n_edges, f);
  (void)ok;
  std::fclose(f);
  Structure s;
  std::string err;
  std::vector<double> ms;
  for (int r = 0; r < reps; ++r) {
    const auto t0 = std::chrono::steady_clock::now();
    if (!build_structure(g, 0, 1, s, err)) { std::printf("build failed: %s\n", err.c_str()); return 1; }
    ms.push_back(std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
  }
  std::sort(ms.begin(), ms.end());
  std::printf("build_structure: median %.3f ms, min %.3f ms over %d (units %d, chunks %d, pairs %d)\n", ms[ms.size() / 2], ms[0], reps, s.n_units, s.n_lchunks, s.n_pairs);
  return 0;
}
