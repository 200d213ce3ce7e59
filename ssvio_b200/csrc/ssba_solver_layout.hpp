// ssba_solver_layout.hpp — shared-memory budget of k_reduced_solve, shared by the host program
// builder (slot allocation) and the kernel launch so that both agree on the layout.
#pragma once

#include <cstddef>

namespace ssba {

constexpr int kSolveThreads = 512;
constexpr int kSolveMaxCols = 64;             // columns per level
constexpr size_t kSolveMaxDynSmem = 226 * 1024;  // opt-in dynamic shared memory of the kernel
constexpr int kSolveMaxCluster = 8;           // CTAs of the thread-block cluster the solver may span

// Cluster size of the reduced solve for a system of n_fp pose columns: one SM is enough for a
// small system (and a __syncthreads is cheaper than a cluster barrier), 4 CTAs below 64 poses, 8
// above; SSBA_SOLVE_CLUSTER overrides (1, 2, 4 or 8).  Used by the host program builder (which deals the
// columns of every level over the CTAs) and recorded in the structure for the launch.
int solver_cluster_size(int n_fp);
// Upper bound from the device (ssba_create asks it once: cudaOccupancyMaxActiveClusters).
void set_solver_cluster_cap(int cap);

struct SolverSmemLayout {
  int x_in_smem;    // y / x vector in shared memory
  int staged;       // level programs double-buffered in shared memory
  int n_slots;      // 6x6 block slots of the factor cache
  size_t off_flag, off_x, off_prog, off_slots, bytes;  // byte offsets into dynamic smem
};

inline SolverSmemLayout solver_smem_layout(int n_fp, int prog_max_seg_ints) {
  SolverSmemLayout L{};
  size_t b = 36 * sizeof(double) * kSolveMaxCols;  // inverse diagonal blocks of the level
  L.off_flag = b;
  b += sizeof(int) * kSolveMaxCols;
  L.off_x = b;
  const size_t xbytes = 6 * sizeof(double) * (size_t)n_fp;
  L.x_in_smem = b + xbytes <= kSolveMaxDynSmem / 4 ? 1 : 0;
  if (L.x_in_smem) b += xbytes;
  b = (b + 15) & ~size_t(15);
  L.off_prog = b;
  const size_t pbytes = 2 * sizeof(int) * (size_t)((prog_max_seg_ints + 3) & ~3);
  L.staged = b + pbytes <= kSolveMaxDynSmem / 2 ? 1 : 0;
  if (L.staged) b += pbytes;
  b = (b + 15) & ~size_t(15);
  L.off_slots = b;
  L.n_slots = (int)((kSolveMaxDynSmem - b) / (36 * sizeof(double)));
  if (L.n_slots < 0) L.n_slots = 0;
  L.bytes = b + (size_t)L.n_slots * 36 * sizeof(double);
  return L;
}

}  // namespace ssba
