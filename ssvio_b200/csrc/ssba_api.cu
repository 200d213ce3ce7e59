// ssba_api.cu — the C ABI of include/ssba.h: handle life cycle, graph upload, the LM driver that
// enqueues trial slots on a CUDA stream, result read-back, multi-GPU plumbing (NCCL, loaded with
// dlopen only when world_size > 1).  No CPU compute path exists here: without a CUDA device
// ssba_create() fails.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "ssba.h"
#include "ssba_device.hpp"
#include "ssba_solver_layout.hpp"
#include "ssba_structure.hpp"

using namespace ssba;

namespace {

thread_local std::string g_create_error;

using Clock = std::chrono::steady_clock;
inline double secs(Clock::time_point a, Clock::time_point b) {
  return std::chrono::duration<double>(b - a).count();
}

// ---- NCCL, resolved at run time (torch's bundled libnccl.so.2 when it is already in the
// process, the system one otherwise); only the handful of entry points the path needs.
struct Nccl {
  void *lib = nullptr;
  typedef struct { char internal[128]; } UniqueId;
  int (*GetUniqueId)(UniqueId *) = nullptr;
  int (*CommInitRank)(void **, int, UniqueId, int) = nullptr;
  int (*CommDestroy)(void *) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  bool load(std::string &err) {
    if (lib) return true;
    lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) { err = std::string("cannot load libnccl: ") + dlerror(); return false; }
    GetUniqueId = (decltype(GetUniqueId))dlsym(lib, "ncclGetUniqueId");
    CommInitRank = (decltype(CommInitRank))dlsym(lib, "ncclCommInitRank");
    CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
    AllReduce = (decltype(AllReduce))dlsym(lib, "ncclAllReduce");
    AllGather = (decltype(AllGather))dlsym(lib, "ncclAllGather");
    GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
    if (!GetUniqueId || !CommInitRank || !CommDestroy || !AllReduce) { err = "libnccl lacks required symbols"; return false; }
    return true;
  }
};
Nccl g_nccl;
constexpr int kNcclDouble = 8;  // ncclFloat64
constexpr int kNcclSum = 0, kNcclMax = 2;

inline size_t align_up(size_t x) { return (x + 255) & ~size_t(255); }

}  // namespace

struct ssba_handle {
  ssba_options opt{};
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  std::string error;
  HostGraph g;
  Structure s;
  bool dirty = true;        // graph changed since the last ssba_initialize
  bool topo_dirty = true;   // ... and not only its VALUES (estimates, measurements, information, Huber widths,
                            // camera parameters): the structure has to be rebuilt
  bool initialized = false;
  // where the value arrays of the resident structure sit in region A (arena and pinned mirror alike): a graph
  // with the same topology is re-uploaded without a structure build (the <= 5 rounds of backend.cpp:175-203)
  size_t off_pose0 = 0, off_point0 = 0;
  long long n_structure_builds = 0, n_structure_reuses = 0;
  double ms_structure_build = 0.0, ms_symbolic = 0.0;
  // SparseOptimizer::setForceStopFlag: raised from any thread, posted to the device on a stream of its own
  std::atomic<int> stop_requested{0};
  cudaStream_t aux_stream = nullptr;
  int *h_one = nullptr;  // pinned constant 1
  // memory: one device arena (static index data first, then work buffers) + pinned mirror of
  // the static part so a whole graph goes up in one copy
  char *d_arena = nullptr; size_t d_arena_bytes = 0;
  char *h_stage = nullptr; size_t h_stage_bytes = 0;      // pinned mirror of region A (big per-edge / per-pair arrays)
  char *h_stage_b = nullptr; size_t h_stage_b_bytes = 0;  // ... of region B (solver program, small index lists)
  Control *h_ctl = nullptr;   // pinned
  double *h_small = nullptr;  // pinned scratch (8 doubles)
  DeviceProblem P{};
  int cur = 0;
  int lin = 0; bool lin_valid = false; double current_chi = 0.0;  // see Control::lin
  double lambda = -1.0, ni = 2.0;
  size_t device_bytes = 0;
  void *comm = nullptr;
  // profiling
  ssba_profile prof{};
  std::vector<cudaEvent_t> ev;
  size_t ev_used = 0;
  struct Span { size_t a, b; int phase; };
  std::vector<Span> spans;
  double setup_seconds = 0.0;
  std::vector<uint8_t> owner_mask;
  // several GPUs: this rank's exchange buffer (cudaMalloc, shared with the peers through CUDA IPC) and the
  // peers' buffers as mapped here; capacity in doubles per partial reduced system
  char *xchg = nullptr; size_t xchg_cap_doubles = 0;
  char *peer[SSBA_MAX_PEERS] = {nullptr};
  bool use_p2p = false;
  long long trial_seq = 0;
  // pose-graph optimisation: its own solver structure and grow-only device buffer
  Structure pg_s;
  char *d_pg = nullptr; size_t d_pg_bytes = 0;
  // pose-only LM: its own grow-only device buffer and pinned staging
  // the edges' values (uv | information | Huber widths) in the caller's order: pinned copy of the caller's arrays
  // (set_edges) and its image on the device; k_gather_edge_values sorts them into the structure's order
  char *h_raw = nullptr; size_t h_raw_bytes = 0;
  char *d_raw = nullptr; size_t d_raw_bytes = 0;
  size_t raw_off_info = 0, raw_off_delta = 0;
  char *d_xr = nullptr; size_t d_xr_bytes = 0;  // device scratch of the pre-sharded structure build's exchanges
  char *d_po = nullptr; size_t d_po_bytes = 0;
  char *h_po = nullptr; size_t h_po_bytes = 0;
};

namespace {

#define CUDA_TRY(h, expr)                                                                   \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      (h)->error = std::string(#expr) + ": " + cudaGetErrorString(_e);                      \
      return SSBA_ERR_CUDA;                                                                 \
    }                                                                                       \
  } while (0)

// programmatic dependent launch between the kernels of a trial: on, unless SSBA_PDL=0 or the per-phase events of
// the profiling mode sit between them
bool pdl_enabled(const ssba_handle *h) {
  static const bool env_on = [] { const char *e = std::getenv("SSBA_PDL"); return !(e && std::atoi(e) == 0); }();
  return env_on && !h->opt.profile;
}

ssba_status fail(ssba_handle *h, ssba_status st, const std::string &msg) {
  if (h) h->error = msg; else g_create_error = msg;
  return st;
}

ssba_status nccl_allreduce(ssba_handle *h, double *buf, size_t n, int op) {
  int rc = g_nccl.AllReduce(buf, buf, n, kNcclDouble, op, h->comm, h->stream);
  if (rc != 0) return fail(h, SSBA_ERR_NCCL, std::string("ncclAllReduce: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error"));
  h->prof.kernel_launches += 0;
  return SSBA_OK;
}

cudaEvent_t next_event(ssba_handle *h) {
  if (h->ev_used == h->ev.size()) {
    cudaEvent_t e; cudaEventCreate(&e); h->ev.push_back(e);
  }
  return h->ev[h->ev_used++];
}

struct PhaseTimer {  // records a CUDA-event span on the launch stream when profiling is on
  ssba_handle *h; int phase; size_t a = 0;
  PhaseTimer(ssba_handle *h_, int phase_) : h(h_), phase(phase_) {
    if (h->opt.profile) { a = h->ev_used; cudaEventRecord(next_event(h), h->stream); }
  }
  ~PhaseTimer() {
    if (h->opt.profile) { size_t b = h->ev_used; cudaEventRecord(next_event(h), h->stream); h->spans.push_back({a, b, phase}); }
  }
};

void collect_profile(ssba_handle *h) {
  if (!h->opt.profile) return;
  for (auto &sp : h->spans) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->ev[sp.a], h->ev[sp.b]);
    switch (sp.phase) {
      case 0: h->prof.ms_linearize += ms; h->prof.n_linearize++; break;
      case 1: h->prof.ms_schur += ms; h->prof.n_schur++; break;
      case 2: h->prof.ms_reduced_solve += ms; h->prof.n_reduced_solve++; break;
      case 3: h->prof.ms_update_chi2 += ms; h->prof.n_update_chi2++; break;
      case 4: h->prof.ms_allreduce += ms; h->prof.n_allreduce++; break;
    }
  }
  h->spans.clear();
  h->ev_used = 0;
}

// ---- peer-memory exchange (several GPUs of one node): every rank allocates an exchange buffer, the CUDA IPC
// handles travel through one ncclAllGather, every rank maps the others' buffers.  Collective: all ranks call
// it with the same `sys_doubles` (they build the same structure).  On any failure the handle simply keeps
// the NCCL path (use_p2p = false) - on every rank, because the outcome is agreed by an all-reduce (min).
// Collective when a buffer exists: every rank unmaps its peers' buffers, then all ranks meet (a one-element
// all-reduce) before anyone frees memory that another process may still have mapped.
void close_peer_exchange(ssba_handle *h) {
  const bool had = h->xchg != nullptr;
  for (int r = 0; r < SSBA_MAX_PEERS; ++r) {
    if (h->peer[r] && h->peer[r] != h->xchg) cudaIpcCloseMemHandle(h->peer[r]);
    h->peer[r] = nullptr;
  }
  if (had && h->comm && h->use_p2p) {
    double *d_one = (double *)(h->xchg + sizeof(PeerHeader));
    if (g_nccl.AllReduce(d_one, d_one, 1, kNcclDouble, kNcclSum, h->comm, h->stream) == 0) cudaStreamSynchronize(h->stream);
  }
  if (h->xchg) cudaFree(h->xchg);
  h->xchg = nullptr; h->xchg_cap_doubles = 0; h->use_p2p = false;
}

ssba_status setup_peer_exchange(ssba_handle *h, size_t sys_doubles) {
  const int world = h->opt.world_size, rank = h->opt.rank;
  if (world <= 1 || world > SSBA_MAX_PEERS || !g_nccl.AllGather) return SSBA_OK;
  if (const char *e = std::getenv("SSBA_P2P")) if (std::atoi(e) == 0) return SSBA_OK;
  if (sys_doubles <= h->xchg_cap_doubles && h->use_p2p) return SSBA_OK;
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  close_peer_exchange(h);
  const size_t cap = sys_doubles + sys_doubles / 2 + 4096;
  const size_t bytes = sizeof(PeerHeader) + 2 * cap * sizeof(double);
  int ok = 1;
  if (cudaMalloc((void **)&h->xchg, bytes) != cudaSuccess) { cudaGetLastError(); ok = 0; }
  cudaIpcMemHandle_t mine;
  std::memset(&mine, 0, sizeof(mine));
  if (ok && cudaMemset(h->xchg, 0, bytes) != cudaSuccess) ok = 0;
  if (ok && cudaIpcGetMemHandle(&mine, h->xchg) != cudaSuccess) { cudaGetLastError(); ok = 0; }
  // handles of all ranks (device staging: NCCL moves device memory)
  char *d_tmp = nullptr;
  std::vector<cudaIpcMemHandle_t> all(world);
  const size_t hb = sizeof(cudaIpcMemHandle_t);
  CUDA_TRY(h, cudaMalloc((void **)&d_tmp, hb * (world + 1) + 64));
  CUDA_TRY(h, cudaMemcpyAsync(d_tmp + hb * world, &mine, hb, cudaMemcpyHostToDevice, h->stream));
  if (g_nccl.AllGather(d_tmp + hb * world, d_tmp, hb, /*ncclChar*/ 0, h->comm, h->stream) != 0) { cudaFree(d_tmp); return fail(h, SSBA_ERR_NCCL, "ncclAllGather failed"); }
  CUDA_TRY(h, cudaMemcpyAsync(all.data(), d_tmp, hb * world, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  for (int r = 0; r < world && ok; ++r) {
    if (r == rank) { h->peer[r] = h->xchg; continue; }
    void *p = nullptr;
    if (cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
    h->peer[r] = (char *)p;
  }
  // agree on the outcome: min over ranks of `ok`
  double okd = ok ? 1.0 : 0.0, *d_ok = (double *)d_tmp;
  CUDA_TRY(h, cudaMemcpyAsync(d_ok, &okd, sizeof(double), cudaMemcpyHostToDevice, h->stream));
  if (g_nccl.AllReduce(d_ok, d_ok, 1, kNcclDouble, /*ncclMin*/ 3, h->comm, h->stream) != 0) { cudaFree(d_tmp); return fail(h, SSBA_ERR_NCCL, "ncclAllReduce failed"); }
  CUDA_TRY(h, cudaMemcpyAsync(&okd, d_ok, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  cudaFree(d_tmp);
  if (okd < 0.5) { close_peer_exchange(h); return SSBA_OK; }
  h->xchg_cap_doubles = cap;
  h->use_p2p = true;
  return SSBA_OK;
}

// one LM trial, stream-ordered; `first` = first slot of an optimize()/step(0) call; `linearize`: k_linearize is part
// of the slot (always, unless k_update linearises the accepted trials itself: then only where the controller asks)
ssba_status enqueue_slot(ssba_handle *h, bool first, bool linearize) {
  const DeviceProblem &P = h->P;
  cudaStream_t st = h->stream;
  const bool multi = h->opt.world_size > 1;
  ssba_status rc;
  if (linearize) {
    PhaseTimer t(h, 0);
    launch_linearize(P, st);
    h->prof.kernel_launches += 1;
  }
  if (first) {
    {
      PhaseTimer t(h, 0);
      launch_fold(P, st);
      h->prof.kernel_launches += 1;
    }
    if (multi) { PhaseTimer t(h, 4); if ((rc = nccl_allreduce(h, P.diag_buf, 6 * (size_t)P.n_fp, kNcclSum))) return rc; }
    launch_maxdiag(P, st);
    if (multi) { PhaseTimer t(h, 4); if ((rc = nccl_allreduce(h, P.scal + 3, 1, kNcclMax))) return rc; }
    launch_lambda_init(P, st);
    h->prof.kernel_launches += 2;
  }
  {
    PhaseTimer t(h, 1);
    launch_schur(P, first, st);
    h->prof.kernel_launches += P.deterministic ? 2 : 1;
  }
  if (multi) {
    PhaseTimer t(h, 4);
    if (P.use_p2p) { launch_exchange_sys(P, st); h->prof.kernel_launches += 1; }
    else if ((rc = nccl_allreduce(h, P.sys, P.sys_doubles, kNcclSum))) return rc;
  }
  {
    PhaseTimer t(h, 2);
    launch_reduced_solve(P, st);
    h->prof.kernel_launches += 1;
  }
  {
    PhaseTimer t(h, 3);
    launch_update(P, !multi, st);
    h->prof.kernel_launches += 1;
    if (multi && !P.use_p2p) { launch_reduce_partials(P, st); h->prof.kernel_launches += 1; }
  }
  if (multi && P.use_p2p) {
    PhaseTimer t(h, 4);
    launch_control_p2p(P, st);  // partial sums through peer memory, then the decision
    h->prof.kernel_launches += 1;
  } else if (multi) {
    { PhaseTimer t(h, 4); if ((rc = nccl_allreduce(h, P.scal, 3, kNcclSum))) return rc; }
    launch_control(P, st);
    h->prof.kernel_launches += 1;
  }
  return SSBA_OK;
}

// run outer iterations until Control::done; returns with h->h_ctl holding the final controller
ssba_status run_lm(ssba_handle *h, int max_iters, bool iteration0) {
  Control &c = *h->h_ctl;
  std::memset(&c, 0, sizeof(c));
  c.tau = h->opt.tau; c.good_lower = h->opt.good_step_lower_scale; c.good_upper = h->opt.good_step_upper_scale;
  c.user_lambda = h->opt.user_lambda_init; c.max_trials = h->opt.max_trials_after_failure;
  c.lambda = h->lambda; c.ni = h->ni;
  c.cur = h->cur; c.first_iteration = iteration0 ? 1 : 0;
  // a step that continues an optimisation (ssba_step(i > 0)) starts from the linearisation k_update made of the
  // accepted trial; a new optimize() linearises (lambda_0 needs max |H_jj| of that very pass)
  const bool keep_lin = h->lin_valid && !iteration0;
  c.lin = h->lin; c.need_linearize = keep_lin ? 0 : 1; c.need_fold = keep_lin ? 1 : 0; c.current_chi = keep_lin ? h->current_chi : 0.0;
  c.max_iters = max_iters; c.last_result = SSBA_SOLVER_OK;
  c.world = h->opt.world_size; c.rank = h->opt.rank;
  c.trial_seq = h->trial_seq;
  if (h->stop_requested.load() && h->opt.world_size == 1) { c.force_stop = 1; c.done = 1; }  // terminate() before the first iteration
  CUDA_TRY(h, cudaMemcpyAsync(h->P.ctl, &c, sizeof(Control), cudaMemcpyHostToDevice, h->stream));
  // the copy above must have left the pinned buffer before it is reused for the read-back
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  bool first = true;
  int guard = 0;
  const int max_slots = max_iters * (c.max_trials > 0 ? c.max_trials : 1) + 1;
  const bool fused_lin = update_linearizes(h->P);
  bool want_lin = c.need_linearize != 0;
  while (!c.done) {
    // optimistic batch: one slot per outstanding outer iteration (every trial accepted)
    int batch = max_iters - c.outer_iter;
    if (batch < 1) batch = 1;
    for (int i = 0; i < batch; ++i) {
      ssba_status rc = enqueue_slot(h, first, !fused_lin || want_lin);
      if (rc) return rc;
      first = false; want_lin = false;
    }
    CUDA_TRY(h, cudaMemcpyAsync(&c, h->P.ctl, sizeof(Control), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    CUDA_TRY(h, cudaGetLastError());
    if (c.done == 2) {  // paused for a linearising slot (see control_step)
      c.done = 0; want_lin = true;
      CUDA_TRY(h, cudaMemcpyAsync(&h->P.ctl->done, &c.done, sizeof(int), cudaMemcpyHostToDevice, h->stream));
      CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    }
    if (c.done) break;
    guard += batch;
    if (guard > max_slots) return fail(h, SSBA_ERR_STATE, "LM driver did not terminate");
  }
  h->cur = c.cur; h->lambda = c.lambda; h->ni = c.ni;
  h->lin = c.lin; h->lin_valid = c.lin_valid != 0; h->current_chi = c.current_chi;
  h->trial_seq = c.trial_seq;
  h->prof.levenberg_iterations += c.n_trials;
  h->prof.outer_iterations += c.outer_iter;
  if (h->opt.world_size > 1 && std::getenv("SSBA_TIMING"))
    std::fprintf(stderr, "[ssba] rank %d last trial: exchange_sys wait %.2f us, sum %.2f us | gap to control %.2f us | fold %.2f us, scal exchange %.2f us, decision %.2f us\n",
                 h->opt.rank, (c.dbg[1] - c.dbg[0]) * 1e-3, (c.dbg[2] - c.dbg[1]) * 1e-3, (c.dbg[3] - c.dbg[2]) * 1e-3, (c.dbg[4] - c.dbg[3]) * 1e-3,
                 (c.dbg[5] - c.dbg[4]) * 1e-3, (c.dbg[6] - c.dbg[5]) * 1e-3);
  collect_profile(h);
  if (c.comm_timeout) return fail(h, SSBA_ERR_NCCL, "peer-memory exchange: a rank did not publish its part in time");
  return SSBA_OK;
}

ssba_status final_chi2(ssba_handle *h, double threshold, double out[4]) {
  launch_final_chi2(h->P, threshold, h->stream);
  h->prof.kernel_launches += 2;
  if (h->opt.world_size > 1) { ssba_status rc = nccl_allreduce(h, h->P.chi_out, 4, kNcclSum); if (rc) return rc; }
  CUDA_TRY(h, cudaMemcpyAsync(h->h_small, h->P.chi_out, 4 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  for (int i = 0; i < 4; ++i) out[i] = h->h_small[i];
  return SSBA_OK;
}

// fixed flags unchanged? (NULL = none fixed)
bool same_flags(const std::vector<uint8_t> &old, int old_n, const uint8_t *now, int n) {
  if (old_n != n || (int)old.size() != n) return false;
  if (now) return n == 0 || std::memcmp(old.data(), now, (size_t)n) == 0;
  for (uint8_t v : old) if (v) return false;
  return true;
}

}  // namespace

// ================================================================================================

extern "C" {

int32_t ssba_version(void) { return SSBA_VERSION_MAJOR * 100 + SSBA_VERSION_MINOR; }

void ssba_default_options(ssba_options *o) {
  if (!o) return;
  std::memset(o, 0, sizeof(*o));
  o->tau = 1e-5;
  o->good_step_lower_scale = 1. / 3.;
  o->good_step_upper_scale = 2. / 3.;
  o->user_lambda_init = 0.0;
  o->max_trials_after_failure = 10;
  o->jacobian_mode = SSBA_JACOBIAN_ANALYTIC;
  o->device_id = -1;
  o->world_size = 1;
}

const char *ssba_last_error(const ssba_handle *h) { return h ? h->error.c_str() : g_create_error.c_str(); }

ssba_status ssba_nccl_unique_id(uint8_t out[SSBA_NCCL_ID_BYTES]) {
  std::string err;
  if (!g_nccl.load(err)) return fail(nullptr, SSBA_ERR_NCCL, err);
  Nccl::UniqueId id;
  if (g_nccl.GetUniqueId(&id) != 0) return fail(nullptr, SSBA_ERR_NCCL, "ncclGetUniqueId failed");
  std::memcpy(out, id.internal, SSBA_NCCL_ID_BYTES);
  return SSBA_OK;
}

ssba_status ssba_create(const ssba_options *opt, ssba_handle **out) {
  if (!out) return fail(nullptr, SSBA_ERR_INVALID_ARG, "out is NULL");
  *out = nullptr;
  ssba_options o;
  if (opt) o = *opt; else ssba_default_options(&o);
  if (o.world_size < 1 || o.rank < 0 || o.rank >= o.world_size) return fail(nullptr, SSBA_ERR_INVALID_ARG, "bad rank/world_size");
  if (o.max_trials_after_failure < 1) return fail(nullptr, SSBA_ERR_INVALID_ARG, "max_trials_after_failure < 1");
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(nullptr, SSBA_ERR_NO_DEVICE, std::string("no CUDA device: libssba has no CPU path (") +
                                                 (ce != cudaSuccess ? cudaGetErrorString(ce) : "0 devices") + ")");
  }
  std::unique_ptr<ssba_handle> h(new ssba_handle);
  h->opt = o;
  if (o.device_id >= 0) {
    if (o.device_id >= ndev) return fail(nullptr, SSBA_ERR_INVALID_ARG, "device_id out of range");
    h->device = o.device_id;
  } else if (cudaGetDevice(&h->device) != cudaSuccess) {
    return fail(nullptr, SSBA_ERR_CUDA, "cudaGetDevice failed");
  }
  if (cudaSetDevice(h->device) != cudaSuccess) return fail(nullptr, SSBA_ERR_CUDA, "cudaSetDevice failed");
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, h->device) != cudaSuccess) return fail(nullptr, SSBA_ERR_CUDA, "cudaGetDeviceProperties failed");
  if (prop.major < 10) return fail(nullptr, SSBA_ERR_NO_DEVICE, "libssba is built for sm_100a (Blackwell) only");
  {
    // per process: the B200s of a node are alike
    static std::once_flag caps_once;
    static int cluster_cap = 1, tree_cap = 1;
    std::call_once(caps_once, [] { cluster_cap = max_solver_cluster(); tree_cap = max_tree_cluster(); });
    set_solver_cluster_cap(cluster_cap);
    set_tree_cluster_cap(tree_cap);
    set_ranks_on_host(o.world_size);  // one process per GPU of one node
  }
  if (o.stream) { h->stream = (cudaStream_t)o.stream; }
  else {
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) return fail(nullptr, SSBA_ERR_CUDA, "cudaStreamCreate failed");
    h->own_stream = true;
  }
  if (cudaHostAlloc((void **)&h->h_ctl, sizeof(Control), cudaHostAllocDefault) != cudaSuccess ||
      cudaHostAlloc((void **)&h->h_small, 64 * sizeof(double), cudaHostAllocDefault) != cudaSuccess ||
      cudaHostAlloc((void **)&h->h_one, sizeof(int), cudaHostAllocDefault) != cudaSuccess)
    return fail(nullptr, SSBA_ERR_ALLOC, "cudaHostAlloc failed");
  *h->h_one = 1;
  if (cudaStreamCreateWithFlags(&h->aux_stream, cudaStreamNonBlocking) != cudaSuccess) return fail(nullptr, SSBA_ERR_CUDA, "cudaStreamCreate failed");
  if (o.world_size > 1) {
    std::string err;
    if (!g_nccl.load(err)) return fail(nullptr, SSBA_ERR_NCCL, err);
    Nccl::UniqueId id;
    std::memcpy(id.internal, o.nccl_id, SSBA_NCCL_ID_BYTES);
    int rc = g_nccl.CommInitRank(&h->comm, o.world_size, id, o.rank);
    if (rc != 0) return fail(nullptr, SSBA_ERR_NCCL, std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error"));
  }
  *out = h.release();
  return SSBA_OK;
}

void ssba_destroy(ssba_handle *h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  close_peer_exchange(h);
  if (h->comm) g_nccl.CommDestroy(h->comm);
  for (auto e : h->ev) cudaEventDestroy(e);
  if (h->d_arena) cudaFree(h->d_arena);
  if (h->h_stage) cudaFreeHost(h->h_stage);
  if (h->h_stage_b) cudaFreeHost(h->h_stage_b);
  if (h->d_po) cudaFree(h->d_po);
  if (h->d_xr) cudaFree(h->d_xr);
  if (h->d_raw) cudaFree(h->d_raw);
  if (h->h_raw) cudaFreeHost(h->h_raw);
  if (h->d_pg) cudaFree(h->d_pg);
  if (h->h_po) cudaFreeHost(h->h_po);
  if (h->h_ctl) cudaFreeHost(h->h_ctl);
  if (h->h_small) cudaFreeHost(h->h_small);
  if (h->h_one) cudaFreeHost(h->h_one);
  if (h->aux_stream) { cudaStreamSynchronize(h->aux_stream); cudaStreamDestroy(h->aux_stream); }
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

ssba_status ssba_set_cameras(ssba_handle *h, const double K[9], int32_t n_cams, const double *ext_qt) {
  if (!h) return SSBA_ERR_INVALID_ARG;
  if (!K || !ext_qt || n_cams < 1 || n_cams > SSBA_MAX_CAMERAS) return fail(h, SSBA_ERR_INVALID_ARG, "set_cameras: bad arguments");
  if (!h->g.have_cams || h->g.cams.n != n_cams) h->topo_dirty = true;  // edges are validated against the camera count
  std::memcpy(h->g.cams.K, K, 9 * sizeof(double));
  std::memcpy(h->g.cams.ext, ext_qt, 7 * sizeof(double) * n_cams);
  h->g.cams.n = n_cams;
  h->g.have_cams = true;
  h->dirty = true;
  return SSBA_OK;
}

ssba_status ssba_set_poses(ssba_handle *h, int32_t n, const double *qt, const uint8_t *fixed) {
  if (!h) return SSBA_ERR_INVALID_ARG;
  if (n < 0 || (n > 0 && !qt)) return fail(h, SSBA_ERR_INVALID_ARG, "set_poses: bad arguments");
  if (!same_flags(h->g.pose_fixed, h->g.n_poses, fixed, n)) h->topo_dirty = true;
  h->g.n_poses = n;
  h->g.poses.assign(qt, qt + 7 * (size_t)n);
  if (fixed) h->g.pose_fixed.assign(fixed, fixed + n); else h->g.pose_fixed.assign(n, 0);
  h->dirty = true;
  return SSBA_OK;
}

ssba_status ssba_set_points(ssba_handle *h, int32_t n, const double *xyz, const uint8_t *fixed) {
  if (!h) return SSBA_ERR_INVALID_ARG;
  if (n < 0 || (n > 0 && !xyz)) return fail(h, SSBA_ERR_INVALID_ARG, "set_points: bad arguments");
  if (!same_flags(h->g.point_fixed, h->g.n_points, fixed, n)) h->topo_dirty = true;
  h->g.n_points = n;
  h->g.points.assign(xyz, xyz + 3 * (size_t)n);
  if (fixed) h->g.point_fixed.assign(fixed, fixed + n); else h->g.point_fixed.assign(n, 0);
  h->dirty = true;
  return SSBA_OK;
}

ssba_status ssba_set_edges(ssba_handle *h, int32_t n, const int32_t *pose_idx, const int32_t *point_idx,
                           const uint8_t *cam_idx, const double *uv, const double *info,
                           const double *huber_delta, double huber_delta_all) {
  if (!h) return SSBA_ERR_INVALID_ARG;
  if (n < 0 || (n > 0 && (!pose_idx || !point_idx || !uv))) return fail(h, SSBA_ERR_INVALID_ARG, "set_edges: bad arguments");
  HostGraph &g = h->g;
  // same topology as the resident graph (indices, cameras, which optional arrays exist)?  Then only the
  // values are taken and ssba_initialize keeps the structure.
  bool same = !h->topo_dirty && g.n_edges == n && (int)g.e_pose.size() == n && (info != nullptr) == g.has_info &&
              (huber_delta != nullptr) == g.has_delta;
  if (same && n > 0) {
    std::vector<CopyJob> cmp = {{g.e_pose.data(), pose_idx, sizeof(int32_t) * (size_t)n}, {g.e_point.data(), point_idx, sizeof(int32_t) * (size_t)n}};
    if (cam_idx) cmp.push_back({g.e_cam.data(), cam_idx, (size_t)n});
    same = parallel_equal(cmp);
    if (same && !cam_idx) for (uint8_t c : g.e_cam) if (c) { same = false; break; }
  }
  if (!same) h->topo_dirty = true;
  g.n_edges = n;
  // the caller keeps its arrays: copy them (on the host thread pool, several megabytes per window)
  g.e_pose.resize(n); g.e_point.resize(n); g.e_cam.resize(n);
  g.e_uv.clear(); g.e_info.clear(); g.e_delta.clear();
  g.values_on_device = true; g.has_info = info != nullptr; g.has_delta = huber_delta != nullptr;
  // the values go: caller -> pinned copy (the caller's arrays are free again when this returns) -> device, the
  // copy over PCIe running while the host goes on to build the structure
  const size_t b_uv = 16 * (size_t)n, b_info = info ? 24 * (size_t)n : 0, b_delta = huber_delta ? 8 * (size_t)n : 0;
  const size_t raw_bytes = align_up(b_uv) + align_up(b_info) + align_up(b_delta) + 256;
  CUDA_TRY(h, cudaSetDevice(h->device));
  if (raw_bytes > h->h_raw_bytes) {
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    if (h->h_raw) cudaFreeHost(h->h_raw);
    if (h->d_raw) cudaFree(h->d_raw);
    h->h_raw = nullptr; h->d_raw = nullptr; h->h_raw_bytes = h->d_raw_bytes = 0;
    const size_t cap = raw_bytes + raw_bytes / 4;
    if (cudaHostAlloc((void **)&h->h_raw, cap, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return fail(h, SSBA_ERR_ALLOC, "cudaHostAlloc failed"); }
    if (cudaMalloc((void **)&h->d_raw, cap) != cudaSuccess) { cudaGetLastError(); return fail(h, SSBA_ERR_ALLOC, "cudaMalloc failed"); }
    h->h_raw_bytes = h->d_raw_bytes = cap;
  } else {
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));  // an earlier upload may still be reading the pinned copy
  }
  h->raw_off_info = align_up(b_uv); h->raw_off_delta = h->raw_off_info + align_up(b_info);
  std::vector<CopyJob> jobs = {{h->h_raw, uv, b_uv}};
  if (!same) {
    jobs.push_back({g.e_pose.data(), pose_idx, sizeof(int32_t) * (size_t)n});
    jobs.push_back({g.e_point.data(), point_idx, sizeof(int32_t) * (size_t)n});
    if (cam_idx) jobs.push_back({g.e_cam.data(), cam_idx, (size_t)n}); else std::fill(g.e_cam.begin(), g.e_cam.end(), (uint8_t)0);
  }
  if (info) jobs.push_back({h->h_raw + h->raw_off_info, info, b_info});
  if (huber_delta) jobs.push_back({h->h_raw + h->raw_off_delta, huber_delta, b_delta});
  parallel_copy(jobs);
  if (n > 0) CUDA_TRY(h, cudaMemcpyAsync(h->d_raw, h->h_raw, h->raw_off_delta + b_delta, cudaMemcpyHostToDevice, h->stream));
  g.delta_all = huber_delta_all;
  h->dirty = true;
  return SSBA_OK;
}

ssba_status ssba_initialize(ssba_handle *h) {
  if (!h) return SSBA_ERR_INVALID_ARG;
  auto t0 = Clock::now();
  CUDA_TRY(h, cudaSetDevice(h->device));
  HostGraph &g = h->g;
  if ((int)g.pose_fixed.size() != g.n_poses || (int)g.point_fixed.size() != g.n_points)
    return fail(h, SSBA_ERR_STATE, "initialize: vertices not set");
  if (g.n_edges == 0) {
    // SparseOptimizer::initializeOptimization: "Attempt to initialize an empty graph"
    // (sparse_optimizer.cpp:202-205) -> optimize() then returns -1
    h->initialized = false;
    return fail(h, SSBA_ERR_EMPTY, "initialize: empty graph (no edges)");
  }
  std::string err;
  Structure &s = h->s;
  const bool timing = std::getenv("SSBA_TIMING") != nullptr;
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));  // nothing of a previous upload may still read the staging buffers

  bool reuse = !h->topo_dirty && h->initialized;
  if (h->opt.world_size > 1 && h->opt.presharded) {
    // pre-sharded input: the structure build is collective, so either every rank keeps its structure or none does
    if (!h->d_xr) { if (cudaMalloc((void **)&h->d_xr, 4096) != cudaSuccess) { cudaGetLastError(); return fail(h, SSBA_ERR_ALLOC, "cudaMalloc failed"); } h->d_xr_bytes = 4096; }
    double flag = reuse ? 0.0 : 1.0;
    CUDA_TRY(h, cudaMemcpyAsync(h->d_xr, &flag, sizeof(double), cudaMemcpyHostToDevice, h->stream));
    ssba_status arc = nccl_allreduce(h, (double *)h->d_xr, 1, kNcclMax);
    if (arc) return arc;
    CUDA_TRY(h, cudaMemcpyAsync(&flag, h->d_xr, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    reuse = flag < 0.5;
  }
  if (reuse) {
    // ---- same topology as the resident structure: only the values go up (estimates, measurements in the
    // structure's edge order, information / Huber widths, cameras), then the state is reset.  This is what
    // rounds 2..5 of backend.cpp:175-203 and the g2o shim's repeated init() cost.
    DeviceProblem &P = h->P;
    const size_t pb = 56 * (size_t)g.n_poses, lb = 24 * (size_t)g.n_points, ne = (size_t)s.n_edges;
    if (pb) std::memcpy(h->h_stage + h->off_pose0, g.poses.data(), pb);
    if (lb) std::memcpy(h->h_stage + h->off_point0, g.points.data(), lb);
    (void)ne;
    if (pb) CUDA_TRY(h, cudaMemcpyAsync(h->d_arena + h->off_pose0, h->h_stage + h->off_pose0, pb, cudaMemcpyHostToDevice, h->stream));
    if (lb) CUDA_TRY(h, cudaMemcpyAsync(h->d_arena + h->off_point0, h->h_stage + h->off_point0, lb, cudaMemcpyHostToDevice, h->stream));
    launch_gather_edge_values(P, (const double *)h->d_raw, g.has_info ? (const double *)(h->d_raw + h->raw_off_info) : nullptr,
                              g.has_delta ? (const double *)(h->d_raw + h->raw_off_delta) : nullptr, h->stream);
    P.cams = g.cams;
    for (int c = 0; c < g.cams.n; ++c) quat_to_matrix(g.cams.ext[c], P.ext_R[c]);
    P.delta_all = g.delta_all;
    h->dirty = false;
    ++h->n_structure_reuses;
    ssba_status rc = ssba_reset_state(h);  // synchronises the stream: the staging buffer is free again
    if (rc) return rc;
    h->setup_seconds = secs(t0, Clock::now());
    if (timing) std::fprintf(stderr, "[ssba] initialize: same topology, values only: %.3f ms\n", 1e3 * h->setup_seconds);
    return SSBA_OK;
  }

  // ---- the arena is planned and filled in two steps so that the upload of the big per-edge /
  // per-pair arrays (region A, several megabytes over PCIe) runs while the host still builds the
  // solver program and the small index lists (region B):
  //   [ A: static big | work buffers sized by edges / pairs / landmarks | B: static small | late work ]
  struct Item { const void *src; size_t bytes; size_t off; void **dst; };
  std::vector<Item> items_a, items_b, work;
  size_t top = 0, bytes_a = 0;
  DeviceProblem &P = h->P;
  std::memset(&P, 0, sizeof(P));
  auto stat = [&](std::vector<Item> &items, const void *src, size_t bytes, const void **dst) {
    if (bytes == 0) { *dst = nullptr; return; }
    Item it{src, bytes, align_up(top), (void **)dst};
    top = it.off + bytes;
    items.push_back(it);
  };
  auto dyn = [&](size_t bytes, void **dst) {
    Item it{nullptr, bytes, align_up(top), dst};
    top = it.off + (bytes ? bytes : 8);
    work.push_back(it);
  };
  auto grow_arena = [&](size_t want) -> bool {
    if (want <= h->d_arena_bytes) return true;
    if (h->d_arena) cudaFree(h->d_arena);
    h->d_arena = nullptr; h->d_arena_bytes = 0;
    if (cudaMalloc((void **)&h->d_arena, want) != cudaSuccess) { cudaGetLastError(); return false; }
    h->d_arena_bytes = want;
    return true;
  };
  auto grow_stage = [&](char *&buf, size_t &cap, size_t want) -> bool {
    if (want <= cap) return true;
    if (buf) cudaFreeHost(buf);
    buf = nullptr; cap = 0;
    const size_t sz = want + want / 4;
    if (cudaHostAlloc((void **)&buf, sz, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return false; }
    cap = sz;
    return true;
  };
#define STAT(items, vec, field) stat(items, (vec).data(), (vec).size() * sizeof((vec)[0]), (const void **)&P.field)
#define DYN(field, count, type) dyn((size_t)(count) * sizeof(type), (void **)&P.field)
  ssba_status cb_rc = SSBA_OK;
  double t_stage_a = 0.0;
  const std::function<void()> on_edges_ready = [&]() {
    auto t_s = Clock::now();
    STAT(items_a, g.poses, pose0); STAT(items_a, g.points, point0);
    STAT(items_a, s.slot_vertex, slot_vertex); STAT(items_a, s.slot_free, slot_free); STAT(items_a, s.slot_pair_ptr, slot_pair_ptr);
    STAT(items_a, s.pair_vertex, pair_vertex); STAT(items_a, s.pair_q, pair_q); STAT(items_a, s.pair_edge_ptr, pair_edge_ptr);
    STAT(items_a, s.pair_slot, pair_slot);
    STAT(items_a, s.e_cam, e_cam);
    STAT(items_a, s.e_orig, e_orig);
    bytes_a = align_up(top);
    DYN(pose[0], 7 * g.n_poses, double); DYN(pose[1], 7 * g.n_poses, double);
    DYN(point[0], 3 * (size_t)g.n_points, double); DYN(point[1], 3 * (size_t)g.n_points, double);
    for (int k = 0; k < 2; ++k) { DYN(W[k], 18 * (size_t)s.n_pairs, double); DYN(Hll[k], 6 * (size_t)s.n_slots, double); DYN(bl[k], 3 * (size_t)s.n_slots, double); }
    DYN(Dinv, 6 * (size_t)s.n_slots, double);
    DYN(pair_rec, (size_t)s.n_pairs, PairRec);  // packed on the device (k_pack_pairs)
    DYN(e_uv, 2 * (size_t)s.n_edges, double);  // filled on the device (k_gather_edge_values)
    if (g.has_info) DYN(e_info, 3 * (size_t)s.n_edges, double);
    if (g.has_delta) DYN(e_delta, (size_t)s.n_edges, double);
    DYN(gather, h->opt.world_size > 1 ? 3 * (size_t)g.n_points : 0, double); DYN(err_out, 2 * (size_t)g.n_edges, double);
    DYN(mask_out, (size_t)g.n_edges, uint8_t);
    const size_t known = align_up(top);
    // room for region B and the late work buffers (a few hundred KB for a sliding window); if the
    // guess is short the arena is re-made below and region A sent again
    if (!grow_arena(known + known / 8 + (4u << 20))) { cb_rc = fail(h, SSBA_ERR_ALLOC, "cudaMalloc failed"); return; }
    if (!grow_stage(h->h_stage, h->h_stage_bytes, bytes_a)) { cb_rc = fail(h, SSBA_ERR_ALLOC, "cudaHostAlloc failed"); return; }
    std::vector<CopyJob> jobs;
    for (auto &it : items_a) jobs.push_back({h->h_stage + it.off, it.src, it.bytes});
    parallel_copy(jobs);
    if (cudaMemcpyAsync(h->d_arena, h->h_stage, bytes_a, cudaMemcpyHostToDevice, h->stream) != cudaSuccess) {
      cb_rc = fail(h, SSBA_ERR_CUDA, "cudaMemcpyAsync (region A) failed");
      return;
    }
    t_stage_a = secs(t_s, Clock::now());
  };
  auto t_a = Clock::now();
  h->initialized = false;  // the device problem is being rebuilt
  h->topo_dirty = true;
  // pre-sharded input (options.presharded, several ranks): what all ranks must agree on travels through NCCL
  // (element-wise maximum of a byte array, sum of a few counters), staged through the pinned / device scratch
  const bool presharded = h->opt.world_size > 1 && h->opt.presharded != 0;
  const AcrossRanks across_ranks = [&](uint8_t *bytes, size_t nb, long long *sums, int ns) -> bool {
    const size_t need = nb + 64;
    if (need > h->d_xr_bytes) {
      if (h->d_xr) cudaFree(h->d_xr);
      h->d_xr = nullptr; h->d_xr_bytes = 0;
      if (cudaMalloc((void **)&h->d_xr, need + need / 2) != cudaSuccess) { cudaGetLastError(); return false; }
      h->d_xr_bytes = need + need / 2;
    }
    double *d_s = (double *)h->d_xr;          // up to 8 counters first (8-byte aligned), then the bytes
    uint8_t *d_b = (uint8_t *)h->d_xr + 64;
    double hs[8] = {0};
    for (int i = 0; i < ns && i < 8; ++i) hs[i] = (double)sums[i];
    if (nb && cudaMemcpyAsync(d_b, bytes, nb, cudaMemcpyHostToDevice, h->stream) != cudaSuccess) return false;
    if (ns && cudaMemcpyAsync(d_s, hs, sizeof(double) * ns, cudaMemcpyHostToDevice, h->stream) != cudaSuccess) return false;
    if (nb && g_nccl.AllReduce(d_b, d_b, nb, /*ncclUint8*/ 1, kNcclMax, h->comm, h->stream) != 0) return false;
    if (ns && g_nccl.AllReduce(d_s, d_s, (size_t)ns, kNcclDouble, kNcclSum, h->comm, h->stream) != 0) return false;
    if (nb && cudaMemcpyAsync(bytes, d_b, nb, cudaMemcpyDeviceToHost, h->stream) != cudaSuccess) return false;
    if (ns && cudaMemcpyAsync(hs, d_s, sizeof(double) * ns, cudaMemcpyDeviceToHost, h->stream) != cudaSuccess) return false;
    if (cudaStreamSynchronize(h->stream) != cudaSuccess) return false;
    for (int i = 0; i < ns && i < 8; ++i) sums[i] = (long long)(hs[i] + 0.5);
    return true;
  };
  if (!build_structure(g, h->opt.rank, h->opt.world_size, s, err, &on_edges_ready, presharded ? &across_ranks : nullptr)) return fail(h, SSBA_ERR_INVALID_ARG, err);
  auto t_b = Clock::now();
  h->ms_structure_build = 1e3 * secs(t_a, t_b);
  h->ms_symbolic = 1e3 * s.seconds_symbolic;
  if (cb_rc) return cb_rc;
  if (s.n_fp + s.n_fl_global == 0) { h->initialized = false; return fail(h, SSBA_ERR_EMPTY, "initialize: 0 vertices to optimize"); }

  // landmarks whose estimate this rank reports in ssba_get_points (world_size > 1)
  h->owner_mask.assign(g.n_points, 0);
  for (int v : s.slot_vertex) h->owner_mask[v] = 1;
  if (h->opt.rank == 0)
    for (int v = 0; v < g.n_points; ++v) if (!s.point_active[v]) h->owner_mask[v] = 1;

  // ---- region B and the late work buffers
  const size_t off_b = align_up(top);
  STAT(items_b, s.lchunk_slot, lchunk_slot);
  STAT(items_b, s.lchunk_lp_ptr, lchunk_lp_ptr); STAT(items_b, s.lp_pair_ptr, lp_pair_ptr); STAT(items_b, s.lp_pair, lp_pair);
  STAT(items_b, s.q_part_ptr, q_part_ptr); STAT(items_b, s.q_part, q_part);
  STAT(items_b, s.pose_of_q, pose_of_q);
  STAT(items_b, h->owner_mask, owner_mask);
  STAT(items_b, s.unit_slot, unit_slot); STAT(items_b, s.unit_n, unit_n); STAT(items_b, s.unit_k, unit_k); STAT(items_b, s.unit_c0, unit_c0);
  STAT(items_b, s.unit_combo_ptr, unit_combo_ptr); STAT(items_b, s.combo_blk, combo_blk);
  STAT(items_b, s.blk_prod_ptr, blk_prod_ptr); STAT(items_b, s.combo_pos, combo_pos);
  STAT(items_b, s.blk_row, blk_row); STAT(items_b, s.blk_col, blk_col); STAT(items_b, s.col_ptr, col_diag);
  STAT(items_b, s.prog, prog); STAT(items_b, s.prog_ptr, prog_ptr);
  const int32_t *d_tree_prog = nullptr;
  stat(items_b, s.tree.words.data(), s.tree.words.size() * sizeof(int32_t), (const void **)&d_tree_prog);
  const size_t bytes_b = align_up(top) - off_b;
  P.n_fin_blocks = std::max(1, (s.n_slots + kReadoutThreads - 1) / kReadoutThreads);
  P.n_lin_blocks = P.n_upd_blocks = s.n_lchunks;
  const int nblk = std::max(P.n_fin_blocks, s.n_lchunks);
  P.sys_doubles = 36 * (size_t)s.n_blocks + 12 * (size_t)s.n_fp;
  DYN(hpp_part[0], 27 * (size_t)s.n_hpp_parts, double); DYN(hpp_part[1], 27 * (size_t)s.n_hpp_parts, double);
  DYN(hpp_fold, 27 * (size_t)s.n_fp, double);
  DYN(sys, P.sys_doubles, double); DYN(xp, 6 * (size_t)s.n_fp, double); DYN(diag_buf, 6 * (size_t)s.n_fp, double);
  DYN(chi_cur_part, nblk, double); DYN(maxdiag_part, nblk, double); DYN(chi_new_part, nblk, double); DYN(scale_part, nblk, double);
  DYN(scal, 8, double); DYN(chi_out, 8, double);
  DYN(stage, 36 * s.combo_blk.size(), double); DYN(stage_b, 6 * s.combo_blk.size(), double);
  DYN(ctl, 1, Control);
  double *d_tree_xchg = nullptr;
  dyn(sizeof(double) * (size_t)std::max(s.tree.xchg_doubles, 2), (void **)&d_tree_xchg);
#undef STAT
#undef DYN
  const size_t total = align_up(top);
  const size_t static_bytes = bytes_a + bytes_b;
  if (total > h->d_arena_bytes) {  // the guess was short: a larger arena, region A goes up again
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    if (!grow_arena(total + total / 4)) return fail(h, SSBA_ERR_ALLOC, "cudaMalloc failed");
    CUDA_TRY(h, cudaMemcpyAsync(h->d_arena, h->h_stage, bytes_a, cudaMemcpyHostToDevice, h->stream));
  }
  if (!grow_stage(h->h_stage_b, h->h_stage_b_bytes, bytes_b)) return fail(h, SSBA_ERR_ALLOC, "cudaHostAlloc failed");
  auto t_c = Clock::now();
  {
    std::vector<CopyJob> jobs;
    for (auto &it : items_b) jobs.push_back({h->h_stage_b + (it.off - off_b), it.src, it.bytes});
    parallel_copy(jobs);
  }
  auto t_d = Clock::now();
  for (auto *v : {&items_a, &items_b, &work}) for (auto &it : *v) *it.dst = h->d_arena + it.off;
  for (auto &it : items_a) {
    if (it.src == (const void *)g.poses.data()) h->off_pose0 = it.off;
    else if (it.src == (const void *)g.points.data()) h->off_point0 = it.off;
  }
  h->device_bytes = total;
  if (bytes_b) CUDA_TRY(h, cudaMemcpyAsync(h->d_arena + off_b, h->h_stage_b, bytes_b, cudaMemcpyHostToDevice, h->stream));

  P.cams = g.cams;
  for (int c = 0; c < g.cams.n; ++c) quat_to_matrix(g.cams.ext[c], P.ext_R[c]);
  P.jacobian_mode = h->opt.jacobian_mode;
  P.delta_all = g.delta_all;
  P.n_poses = g.n_poses; P.n_points = g.n_points; P.n_fp = s.n_fp; P.n_slots = s.n_slots; P.n_pairs = s.n_pairs;
  P.n_edges = s.n_edges; P.n_blocks = s.n_blocks; P.n_hpp_parts = s.n_hpp_parts; P.n_levels = s.n_levels;
  P.n_edges_total = g.n_edges;
  P.prog_max_seg = s.prog_max_seg;
  P.n_segments = s.n_segments;
  P.solve_cluster = s.solve_cluster;
  fill_tree_dev(s.tree, d_tree_prog, d_tree_xchg, P.tree);
  if (h->opt.world_size > 1) {
    ssba_status prc = setup_peer_exchange(h, P.sys_doubles);
    if (prc) return prc;
    P.use_p2p = h->use_p2p ? 1 : 0;
    for (int r = 0; r < SSBA_MAX_PEERS; ++r) P.peer[r] = h->peer[r];
    // both partial systems of this graph's layout start out zero (no peer can still be reading: it
    // published the flags the last k_control of this rank waited for only after it was done)
    if (h->use_p2p) CUDA_TRY(h, cudaMemsetAsync(h->xchg + sizeof(PeerHeader), 0, 2 * sizeof(double) * P.sys_doubles, h->stream));
  }
  P.n_units = s.n_units;
  launch_pack_pairs(P, h->stream);
  launch_gather_edge_values(P, (const double *)h->d_raw, g.has_info ? (const double *)(h->d_raw + h->raw_off_info) : nullptr,
                            g.has_delta ? (const double *)(h->d_raw + h->raw_off_delta) : nullptr, h->stream);
  P.pdl = pdl_enabled(h) ? 1 : 0;
  {
    static const bool stage = [] { const char *e = std::getenv("SSBA_POSE_STAGE"); return !(e && std::atoi(e) == 0); }();
    P.pose_stage = stage ? pose_stage_bytes(P.n_poses) : 0u;
  }
  {
    static const bool det = [] { const char *e = std::getenv("SSBA_DETERMINISTIC"); return !(e && std::atoi(e) == 0); }();
    P.deterministic = det ? 1 : 0;
  }
  h->initialized = true;
  h->dirty = false;
  h->topo_dirty = false;
  ++h->n_structure_builds;
  ssba_status rc = ssba_reset_state(h);
  if (rc) return rc;
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));  // h_stage may be rewritten by the next initialize
  h->setup_seconds = secs(t0, Clock::now());
  if (timing)
    std::fprintf(stderr, "[ssba] initialize: build_structure %.3f ms (of which staging + upload start of region A %.3f ms, %.1f MB), "
                 "plan %.3f ms, stage B %.3f ms (%.2f MB), upload wait+reset+sync %.3f ms, total %.3f ms\n", 1e3 * secs(t_a, t_b),
                 1e3 * t_stage_a, bytes_a / 1e6, 1e3 * secs(t_b, t_c), 1e3 * secs(t_c, t_d), bytes_b / 1e6,
                 1e3 * secs(t_d, Clock::now()), 1e3 * h->setup_seconds);
  (void)static_bytes;
  return SSBA_OK;
}

ssba_status ssba_drop_structure(ssba_handle *h) {
  if (!h) return SSBA_ERR_INVALID_ARG;
  h->topo_dirty = true;
  h->dirty = true;
  return SSBA_OK;
}

ssba_status ssba_reset_state(ssba_handle *h) {
  if (!h) return SSBA_ERR_INVALID_ARG;
  if (!h->initialized) return fail(h, SSBA_ERR_STATE, "reset_state: not initialised");
  CUDA_TRY(h, cudaSetDevice(h->device));
  const DeviceProblem &P = h->P;
  const size_t pb = 7 * sizeof(double) * (size_t)P.n_poses, lb = 3 * sizeof(double) * (size_t)P.n_points;
  for (int k = 0; k < 2; ++k) {
    if (pb) CUDA_TRY(h, cudaMemcpyAsync(P.pose[k], P.pose0, pb, cudaMemcpyDeviceToDevice, h->stream));
    if (lb) CUDA_TRY(h, cudaMemcpyAsync(P.point[k], P.point0, lb, cudaMemcpyDeviceToDevice, h->stream));
  }
  // k_schur accumulates into a zeroed reduced system; after the first trial k_update keeps it so
  CUDA_TRY(h, cudaMemsetAsync(P.sys, 0, sizeof(double) * P.sys_doubles, h->stream));
  h->cur = 0; h->lambda = -1.0; h->ni = 2.0;
  h->lin = 0; h->lin_valid = false;
  // the controller must name buffer 0 for the read-out kernels even before the first optimize
  std::memset(h->h_ctl, 0, sizeof(Control));
  h->h_ctl->world = h->opt.world_size; h->h_ctl->rank = h->opt.rank;
  CUDA_TRY(h, cudaMemcpyAsync(P.ctl, h->h_ctl, sizeof(Control), cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  return SSBA_OK;
}

static ssba_status ensure_ready(ssba_handle *h) {
  if (h->dirty || !h->initialized) return ssba_initialize(h);
  CUDA_TRY(h, cudaSetDevice(h->device));
  return SSBA_OK;
}

ssba_status ssba_optimize(ssba_handle *h, int32_t max_iters, ssba_report *report) {
  if (!h) return SSBA_ERR_INVALID_ARG;
  auto t0 = Clock::now();
  if (report) std::memset(report, 0, sizeof(*report));
  if (max_iters > SSBA_MAX_ITER_RECORDS) return fail(h, SSBA_ERR_INVALID_ARG, "optimize: max_iters too large");
  h->setup_seconds = 0.0;
  ssba_status rc = ensure_ready(h);
  if (rc == SSBA_ERR_EMPTY) { if (report) { report->iterations = -1; report->last_result = SSBA_SOLVER_FAIL; } return rc; }
  if (rc) return rc;
  if (max_iters <= 0) {  // optimize(0): the loop body never runs (sparse_optimizer.cpp:387)
    if (report) { report->iterations = 0; report->last_result = SSBA_SOLVER_OK; }
    return SSBA_OK;
  }
  rc = run_lm(h, max_iters, true);
  if (rc) return rc;
  if (report) {
    const Control &c = *h->h_ctl;
    double f[4];
    rc = final_chi2(h, 0.0, f);
    if (rc) return rc;
    report->iterations = c.outer_iter;
    report->last_result = c.last_result;
    report->n_records = c.n_records;
    report->cholesky_failures = c.cholesky_failures;
    report->chi2_initial = c.chi2_initial;
    report->chi2_plain = f[0];
    report->chi2_robust = f[1];
    report->lambda = c.lambda;
    std::memcpy(report->iters, c.records, sizeof(ssba_iter_record) * c.n_records);
    report->seconds_setup = h->setup_seconds;
    report->seconds_total = secs(t0, Clock::now());
  }
  return SSBA_OK;
}

ssba_status ssba_step(ssba_handle *h, int32_t iteration, int32_t *solver_result, ssba_iter_record *record) {
  if (!h) return SSBA_ERR_INVALID_ARG;
  ssba_status rc = ensure_ready(h);
  if (rc) return rc;
  if (iteration > 0 && h->lambda < 0) return fail(h, SSBA_ERR_STATE, "step: iteration > 0 before iteration 0");
  rc = run_lm(h, 1, iteration == 0);
  if (rc) return rc;
  const Control &c = *h->h_ctl;
  if (solver_result) *solver_result = c.last_result;
  if (record && c.n_records > 0) *record = c.records[c.n_records - 1];
  return SSBA_OK;
}

// Device -> caller's (pageable) buffer through the handle's pinned staging buffer: one DMA at PCIe speed, then a host
// copy (threaded when large), instead of the driver's chunked staging of a pageable destination.  The staging buffer
// is idle whenever a read-out can run (ssba_initialize waits for its upload).
static ssba_status d2h(ssba_handle *h, void *out, const void *src, size_t bytes) {
  if (bytes == 0) return SSBA_OK;
  if (h->h_stage && bytes <= h->h_stage_bytes) {
    CUDA_TRY(h, cudaMemcpyAsync(h->h_stage, src, bytes, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    if (bytes >= (1u << 20)) parallel_copy({{out, h->h_stage, bytes}}); else std::memcpy(out, h->h_stage, bytes);
    return SSBA_OK;
  }
  CUDA_TRY(h, cudaMemcpyAsync(out, src, bytes, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  return SSBA_OK;
}

ssba_status ssba_get_poses(ssba_handle *h, double *out) {
  if (!h || !out) return SSBA_ERR_INVALID_ARG;
  if (!h->initialized || h->dirty) {  // nothing ran on the graph as it is now: the estimates are the ones that were set
    if ((int)h->g.poses.size() != 7 * h->g.n_poses) return fail(h, SSBA_ERR_STATE, "get_poses: no poses");
    std::memcpy(out, h->g.poses.data(), h->g.poses.size() * sizeof(double));
    return SSBA_OK;
  }
  CUDA_TRY(h, cudaSetDevice(h->device));
  return d2h(h, out, h->P.pose[h->cur], 7 * sizeof(double) * (size_t)h->P.n_poses);
}

ssba_status ssba_get_points(ssba_handle *h, double *out) {
  if (!h || !out) return SSBA_ERR_INVALID_ARG;
  if (!h->initialized || h->dirty) {
    if ((int)h->g.points.size() != 3 * h->g.n_points) return fail(h, SSBA_ERR_STATE, "get_points: no points");
    std::memcpy(out, h->g.points.data(), h->g.points.size() * sizeof(double));
    return SSBA_OK;
  }
  CUDA_TRY(h, cudaSetDevice(h->device));
  if (h->opt.world_size > 1) {
    // every rank owns the estimates of its landmark shard only: zero the rest, sum over ranks
    // (rank 0 also contributes the landmarks no rank optimises)
    launch_gather_points(h->P, h->stream);
    h->prof.kernel_launches += 1;
    ssba_status rc = nccl_allreduce(h, h->P.gather, 3 * (size_t)h->P.n_points, kNcclSum);
    if (rc) return rc;
    return d2h(h, out, h->P.gather, 3 * sizeof(double) * (size_t)h->P.n_points);
  }
  return d2h(h, out, h->P.point[h->cur], 3 * sizeof(double) * (size_t)h->P.n_points);
}

ssba_status ssba_get_edge_errors(ssba_handle *h, double *out) {
  if (!h || !out) return SSBA_ERR_INVALID_ARG;
  ssba_status rc = ensure_ready(h);
  if (rc) return rc;
  const size_t bytes = 2 * sizeof(double) * (size_t)h->P.n_edges_total;
  CUDA_TRY(h, cudaMemsetAsync(h->P.err_out, 0, bytes, h->stream));
  launch_edge_errors(h->P, h->stream);
  h->prof.kernel_launches += 1;
  // replicated input: every edge lives on exactly one rank, the sum is the union; pre-sharded input: the edges are this rank's own
  if (h->opt.world_size > 1 && !h->opt.presharded) { rc = nccl_allreduce(h, h->P.err_out, 2 * (size_t)h->P.n_edges_total, kNcclSum); if (rc) return rc; }
  return d2h(h, out, h->P.err_out, bytes);
}

ssba_status ssba_chi2(ssba_handle *h, double *plain, double *robust) {
  if (!h) return SSBA_ERR_INVALID_ARG;
  ssba_status rc = ensure_ready(h);
  if (rc) return rc;
  double f[4];
  rc = final_chi2(h, 0.0, f);
  if (rc) return rc;
  if (plain) *plain = f[0];
  if (robust) *robust = f[1];
  return SSBA_OK;
}

ssba_status ssba_count_outliers(ssba_handle *h, double thr, int64_t *n_out, int64_t *n_in) {
  if (!h) return SSBA_ERR_INVALID_ARG;
  ssba_status rc = ensure_ready(h);
  if (rc) return rc;
  double f[4];
  rc = final_chi2(h, thr, f);
  if (rc) return rc;
  if (n_out) *n_out = (int64_t)(f[2] + 0.5);
  // edges whose two vertices are fixed are never active (sparse_optimizer.cpp:237): their
  // _error stays zero in the reference, so backend.cpp:184 counts them as inliers
  if (n_in) *n_in = (int64_t)(f[3] + 0.5) + h->s.n_inactive_edges_global;
  return SSBA_OK;
}

ssba_status ssba_get_outlier_mask(ssba_handle *h, double thr, uint8_t *mask_out, int64_t *n_out) {
  if (!h || !mask_out) return SSBA_ERR_INVALID_ARG;
  ssba_status rc = ensure_ready(h);
  if (rc) return rc;
  const size_t n = (size_t)h->P.n_edges_total;
  CUDA_TRY(h, cudaMemsetAsync(h->P.mask_out, 0, n, h->stream));  // inactive edges (both ends fixed) are never outliers
  launch_outlier_mask(h->P, thr, h->stream);
  h->prof.kernel_launches += 2;
  if (h->opt.world_size > 1) {  // every edge lives on exactly one rank: the sum is the union
    if (!h->opt.presharded && g_nccl.AllReduce(h->P.mask_out, h->P.mask_out, n, /*ncclUint8*/ 1, kNcclSum, h->comm, h->stream) != 0) return fail(h, SSBA_ERR_NCCL, "ncclAllReduce failed");
    if ((rc = nccl_allreduce(h, h->P.chi_out, 4, kNcclSum))) return rc;
  }
  CUDA_TRY(h, cudaMemcpyAsync(mask_out, h->P.mask_out, n, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaMemcpyAsync(h->h_small, h->P.chi_out, 4 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  if (n_out) *n_out = (int64_t)(h->h_small[2] + 0.5);
  return SSBA_OK;
}

ssba_status ssba_request_stop(ssba_handle *h) {
  if (!h) return SSBA_ERR_INVALID_ARG;
  if (h->opt.world_size > 1) return fail(h, SSBA_ERR_STATE, "request_stop: not supported on a multi-rank handle (the ranks must agree on every trial)");
  h->stop_requested.store(1);
  // a running optimize() sees it at its next trial: 4 bytes into the device-resident controller, on a stream of
  // its own (the launch stream is busy with the enqueued trials)
  if (h->initialized && h->P.ctl) {
    if (cudaSetDevice(h->device) != cudaSuccess) return SSBA_ERR_CUDA;
    if (cudaMemcpyAsync(&h->P.ctl->force_stop, h->h_one, sizeof(int), cudaMemcpyHostToDevice, h->aux_stream) != cudaSuccess) { cudaGetLastError(); return SSBA_ERR_CUDA; }
  }
  return SSBA_OK;
}

ssba_status ssba_clear_stop(ssba_handle *h) {
  if (!h) return SSBA_ERR_INVALID_ARG;
  if (h->aux_stream) cudaStreamSynchronize(h->aux_stream);  // a pending raise must not land after this
  h->stop_requested.store(0);
  return SSBA_OK;
}

ssba_status ssba_optimize_rounds(ssba_handle *h, int32_t max_rounds, int32_t iters_per_round,
                                 double chi2_threshold, double inlier_ratio, int32_t *rounds_done,
                                 int64_t *n_outliers, int64_t *n_inliers, ssba_report *last_report) {
  if (!h) return SSBA_ERR_INVALID_ARG;
  if (max_rounds < 1) return fail(h, SSBA_ERR_INVALID_ARG, "optimize_rounds: max_rounds < 1");
  int round = 0;
  int64_t no = 0, ni = 0;
  ssba_report rep;
  while (round < max_rounds) {  // backend.cpp:175
    ssba_status rc = ssba_optimize(h, iters_per_round, &rep);  // :177-178
    if (rc) { if (last_report) *last_report = rep; return rc; }
    ++round;
    rc = ssba_count_outliers(h, chi2_threshold, &no, &ni);  // :179-192
    if (rc) return rc;
    const double ratio = (ni + no) > 0 ? ni / double(ni + no) : 1.0;
    if (ratio > inlier_ratio) break;  // :193-201
  }
  if (rounds_done) *rounds_done = round;
  if (n_outliers) *n_outliers = no;
  if (n_inliers) *n_inliers = ni;
  if (last_report) *last_report = rep;
  return SSBA_OK;
}

ssba_status ssba_plan_shards(int32_t n_poses, const uint8_t *pose_fixed, int32_t n_points,
                             const uint8_t *point_fixed, int32_t n_edges, const int32_t *pose_idx,
                             const int32_t *point_idx, int32_t world_size, int32_t *owner_out) {
  if (n_poses < 0 || n_points < 0 || n_edges < 0 || world_size < 1 || !owner_out ||
      (n_edges > 0 && (!pose_idx || !point_idx)))
    return fail(nullptr, SSBA_ERR_INVALID_ARG, "plan_shards: bad arguments");
  HostGraph g;
  g.have_cams = true; g.cams.n = 1;
  g.n_poses = n_poses; g.n_points = n_points; g.n_edges = n_edges;
  g.poses.assign(7 * (size_t)n_poses, 0.0); g.points.assign(3 * (size_t)n_points, 0.0);
  if (pose_fixed) g.pose_fixed.assign(pose_fixed, pose_fixed + n_poses); else g.pose_fixed.assign(n_poses, 0);
  if (point_fixed) g.point_fixed.assign(point_fixed, point_fixed + n_points); else g.point_fixed.assign(n_points, 0);
  g.e_pose.assign(pose_idx, pose_idx + n_edges); g.e_point.assign(point_idx, point_idx + n_edges);
  g.e_cam.assign(n_edges, 0); g.e_uv.assign(2 * (size_t)n_edges, 0.0);
  std::vector<int32_t> owner;
  std::string err;
  if (!plan_shards(g, world_size, owner, err)) return fail(nullptr, SSBA_ERR_INVALID_ARG, err);
  std::memcpy(owner_out, owner.data(), sizeof(int32_t) * (size_t)n_points);
  return SSBA_OK;
}

static ssba_status pose_only_impl(ssba_handle *h, const double K[9], int32_t n_frames, const int32_t *feat_ptr,
                                  const double *poses_in, const double *xyz, const double *uv, int32_t rounds,
                                  int32_t iters, int32_t pre_rounds, double chi2_threshold, double *poses_out, uint8_t *outlier_out,
                                  int32_t *n_inliers_out, double *chi2_out) {
  if (!h) return SSBA_ERR_INVALID_ARG;
  if (!K || n_frames < 0 || rounds < 0 || iters < 0 || (n_frames > 0 && (!feat_ptr || !poses_in || !poses_out)))
    return fail(h, SSBA_ERR_INVALID_ARG, "pose_only_optimize: bad arguments");
  if (n_frames == 0) return SSBA_OK;
  const int64_t n = feat_ptr[n_frames];
  if (feat_ptr[0] != 0 || n < 0 || (n > 0 && (!xyz || !uv))) return fail(h, SSBA_ERR_INVALID_ARG, "pose_only_optimize: bad feature arrays");
  for (int f = 0; f < n_frames; ++f)
    if (feat_ptr[f + 1] < feat_ptr[f]) return fail(h, SSBA_ERR_INVALID_ARG, "pose_only_optimize: feat_ptr not monotone");
  CUDA_TRY(h, cudaSetDevice(h->device));
  // one device buffer, one pinned mirror: [feat_ptr | poses_in | xyz | uv] up, [poses | chi2 | inliers | flags] down
  size_t top = 0;
  auto place = [&](size_t bytes) { const size_t o = align_up(top); top = o + bytes; return o; };
  const size_t o_fp = place(sizeof(int32_t) * (size_t)(n_frames + 1)), o_pin = place(56 * (size_t)n_frames),
               o_xyz = place(24 * (size_t)n), o_uv = place(16 * (size_t)n);
  const size_t up_bytes = align_up(top);
  const size_t o_pout = place(56 * (size_t)n_frames), o_chi = place(8 * (size_t)n_frames),
               o_nin = place(4 * (size_t)n_frames), o_flag = place((size_t)n);
  const size_t down_end = align_up(top);
  const size_t o_err = place(16 * (size_t)n);
  const size_t total = align_up(top);
  if (total > h->d_po_bytes) {
    if (h->d_po) cudaFree(h->d_po);
    h->d_po = nullptr; h->d_po_bytes = 0;
    if (cudaMalloc((void **)&h->d_po, total + total / 4) != cudaSuccess) { cudaGetLastError(); return fail(h, SSBA_ERR_ALLOC, "cudaMalloc failed"); }
    h->d_po_bytes = total + total / 4;
  }
  if (down_end > h->h_po_bytes) {
    if (h->h_po) cudaFreeHost(h->h_po);
    h->h_po = nullptr; h->h_po_bytes = 0;
    if (cudaHostAlloc((void **)&h->h_po, down_end + down_end / 4, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return fail(h, SSBA_ERR_ALLOC, "cudaHostAlloc failed"); }
    h->h_po_bytes = down_end + down_end / 4;
  }
  std::memcpy(h->h_po + o_fp, feat_ptr, sizeof(int32_t) * (size_t)(n_frames + 1));
  std::memcpy(h->h_po + o_pin, poses_in, 56 * (size_t)n_frames);
  if (n) { std::memcpy(h->h_po + o_xyz, xyz, 24 * (size_t)n); std::memcpy(h->h_po + o_uv, uv, 16 * (size_t)n); }
  CUDA_TRY(h, cudaMemcpyAsync(h->d_po, h->h_po, up_bytes, cudaMemcpyHostToDevice, h->stream));
  char *d = h->d_po;
  launch_pose_only(K, n_frames, rounds, iters, pre_rounds, h->opt.max_trials_after_failure, chi2_threshold, h->opt.tau,
                   h->opt.good_step_lower_scale, h->opt.good_step_upper_scale, h->opt.user_lambda_init, (const int32_t *)(d + o_fp),
                   (const double *)(d + o_pin), (const double *)(d + o_xyz), (const double *)(d + o_uv), (double *)(d + o_err),
                   (uint8_t *)(d + o_flag), (double *)(d + o_pout), (double *)(d + o_chi), (int32_t *)(d + o_nin), h->stream);
  h->prof.kernel_launches += 1;
  CUDA_TRY(h, cudaMemcpyAsync(h->h_po + o_pout, d + o_pout, down_end - o_pout, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  CUDA_TRY(h, cudaGetLastError());
  std::memcpy(poses_out, h->h_po + o_pout, 56 * (size_t)n_frames);
  if (chi2_out) std::memcpy(chi2_out, h->h_po + o_chi, 8 * (size_t)n_frames);
  if (n_inliers_out) std::memcpy(n_inliers_out, h->h_po + o_nin, 4 * (size_t)n_frames);
  if (outlier_out && n) std::memcpy(outlier_out, h->h_po + o_flag, (size_t)n);
  return SSBA_OK;
}

ssba_status ssba_pose_only_optimize(ssba_handle *h, const double K[9], int32_t n_frames, const int32_t *feat_ptr,
                                    const double *poses_in, const double *xyz, const double *uv, int32_t rounds,
                                    int32_t iters, double chi2_threshold, double *poses_out, uint8_t *outlier_out,
                                    int32_t *n_inliers_out, double *chi2_out) {
  return pose_only_impl(h, K, n_frames, feat_ptr, poses_in, xyz, uv, rounds, iters, 0, chi2_threshold, poses_out, outlier_out,
                        n_inliers_out, chi2_out);
}

ssba_status ssba_pose_only_optimize_loop(ssba_handle *h, const double K[9], int32_t n_frames, const int32_t *feat_ptr,
                                         const double *poses_in, const double *xyz, const double *uv, int32_t rounds,
                                         int32_t iters, double chi2_threshold, double *poses_out, uint8_t *outlier_out,
                                         int32_t *n_inliers_out, double *chi2_out) {
  return pose_only_impl(h, K, n_frames, feat_ptr, poses_in, xyz, uv, rounds, iters, 1, chi2_threshold, poses_out, outlier_out,
                        n_inliers_out, chi2_out);
}

ssba_status ssba_pose_graph_optimize(ssba_handle *h, int32_t n_poses, const double *poses_in, const uint8_t *fixed,
                                     int32_t n_edges, const int32_t *v0, const int32_t *v1, const double *meas,
                                     int32_t iters, double *poses_out, ssba_report *report) {
  if (!h) return SSBA_ERR_INVALID_ARG;
  if (report) std::memset(report, 0, sizeof(*report));
  if (n_poses < 0 || n_edges < 0 || iters < 0 || iters > SSBA_MAX_ITER_RECORDS || (n_poses > 0 && (!poses_in || !poses_out)) ||
      (n_edges > 0 && (!v0 || !v1 || !meas)))
    return fail(h, SSBA_ERR_INVALID_ARG, "pose_graph_optimize: bad arguments");
  auto t0 = Clock::now();
  CUDA_TRY(h, cudaSetDevice(h->device));
  // ---- active edges (not both ends fixed, sparse_optimizer.cpp:237), free key-frames, pattern of H
  std::vector<int32_t> ev0, ev1, fidx(n_poses, -1), free_rows;
  std::vector<double> minv;
  for (int e = 0; e < n_edges; ++e) {
    if ((unsigned)v0[e] >= (unsigned)n_poses || (unsigned)v1[e] >= (unsigned)n_poses || v0[e] == v1[e])
      return fail(h, SSBA_ERR_INVALID_ARG, "pose_graph_optimize: bad edge");
    if (fixed && fixed[v0[e]] && fixed[v1[e]]) continue;
    ev0.push_back(v0[e]); ev1.push_back(v1[e]);
    double mi[7];
    se3_inverse(meas + 7 * (size_t)e, mi);
    minv.insert(minv.end(), mi, mi + 7);
  }
  const int ne = (int)ev0.size();
  {
    std::vector<uint8_t> touched(n_poses, 0);
    for (int e = 0; e < ne; ++e) { touched[ev0[e]] = 1; touched[ev1[e]] = 1; }
    for (int i = 0; i < n_poses; ++i)
      if (touched[i] && !(fixed && fixed[i])) { fidx[i] = (int)free_rows.size(); free_rows.push_back(i); }
  }
  const int nf = (int)free_rows.size();
  if (nf == 0) {
    std::memcpy(poses_out, poses_in, 56 * (size_t)n_poses);
    if (report) { report->iterations = -1; report->last_result = SSBA_SOLVER_FAIL; }
    return fail(h, SSBA_ERR_EMPTY, "pose_graph_optimize: 0 vertices to optimize");
  }
  std::vector<std::vector<int>> adj(nf);
  for (int e = 0; e < ne; ++e) {
    const int a = fidx[ev0[e]], b = fidx[ev1[e]];
    if (a < 0 || b < 0) continue;
    adj[std::min(a, b)].push_back(std::max(a, b));
  }
  for (auto &c : adj) { std::sort(c.begin(), c.end()); c.erase(std::unique(c.begin(), c.end()), c.end()); }
  Structure &s = h->pg_s;
  std::vector<int> perm;
  std::string err;
  if (!build_solver_structure(nf, adj, s, perm, err)) return fail(h, SSBA_ERR_INVALID_ARG, err);
  std::vector<int32_t> q_of_free(nf), pose_of_q(nf), eq0(ne), eq1(ne), eblk(ne, -1);
  for (int q = 0; q < nf; ++q) { q_of_free[perm[q]] = q; pose_of_q[q] = free_rows[perm[q]]; }
  for (int e = 0; e < ne; ++e) {
    eq0[e] = fidx[ev0[e]] >= 0 ? q_of_free[fidx[ev0[e]]] : -1;
    eq1[e] = fidx[ev1[e]] >= 0 ? q_of_free[fidx[ev1[e]]] : -1;
    if (eq0[e] >= 0 && eq1[e] >= 0) {
      const int col = std::min(eq0[e], eq1[e]), row = std::max(eq0[e], eq1[e]);
      const int32_t *b0 = s.blk_row.data() + s.col_ptr[col], *b1 = s.blk_row.data() + s.col_ptr[col + 1];
      const int32_t *it = std::lower_bound(b0, b1, row);
      if (it == b1 || *it != row) return fail(h, SSBA_ERR_STATE, "internal: pose-graph block missing from the factor pattern");
      eblk[e] = (int32_t)(it - s.blk_row.data());
    }
  }
  // producers of every block of H and of every right-hand side, in edge order (the fixed summation order of
  // k_pg_assemble): offsets into the per-edge staging records of k_pg_linearize
  std::vector<int32_t> prod_ptr(s.n_blocks + nf + 1, 0), prod;
  {
    auto lists_of = [&](int e, int out[5][2]) {  // {list, offset in the record} of what edge e produces; list < 0: nothing
      const int q0 = eq0[e], q1 = eq1[e];
      out[0][0] = q0 >= 0 ? s.col_ptr[q0] : -1; out[0][1] = 0;            // diagonal block of q0 (first block of its column)
      out[1][0] = q1 >= 0 ? s.col_ptr[q1] : -1; out[1][1] = 36;
      out[2][0] = eblk[e]; out[2][1] = 72;
      out[3][0] = q0 >= 0 ? s.n_blocks + q0 : -1; out[3][1] = 108;
      out[4][0] = q1 >= 0 ? s.n_blocks + q1 : -1; out[4][1] = 114;
    };
    int tmp[5][2];
    for (int e = 0; e < ne; ++e) { lists_of(e, tmp); for (auto &t : tmp) if (t[0] >= 0) ++prod_ptr[t[0] + 1]; }
    for (size_t i = 1; i < prod_ptr.size(); ++i) prod_ptr[i] += prod_ptr[i - 1];
    prod.resize(prod_ptr.back());
    std::vector<int32_t> fillp(prod_ptr.begin(), prod_ptr.end() - 1);
    for (int e = 0; e < ne; ++e) { lists_of(e, tmp); for (auto &t : tmp) if (t[0] >= 0) prod[fillp[t[0]]++] = 120 * e + t[1]; }
  }
  // ---- device buffer: static part (uploaded), then work buffers
  DeviceProblem P;
  std::memset(&P, 0, sizeof(P));
  P.n_poses = n_poses; P.n_fp = nf; P.n_blocks = s.n_blocks; P.n_levels = s.n_levels;
  P.prog_max_seg = s.prog_max_seg; P.n_segments = s.n_segments; P.solve_cluster = s.solve_cluster;
  P.sys_doubles = 36 * (size_t)s.n_blocks + 12 * (size_t)nf;
  struct Item { const void *src; size_t bytes; size_t off; };
  std::vector<Item> items;
  size_t top = 0;
  auto place = [&](const void *src, size_t bytes) { const size_t o = align_up(top); top = o + (bytes ? bytes : 8); items.push_back({src, bytes, o}); return o; };
  const size_t o_pose0 = place(poses_in, 56 * (size_t)n_poses), o_ev0 = place(ev0.data(), 4 * (size_t)ne), o_ev1 = place(ev1.data(), 4 * (size_t)ne),
               o_eq0 = place(eq0.data(), 4 * (size_t)ne), o_eq1 = place(eq1.data(), 4 * (size_t)ne), o_eblk = place(eblk.data(), 4 * (size_t)ne),
               o_minv = place(minv.data(), 56 * (size_t)ne), o_poq = place(pose_of_q.data(), 4 * (size_t)nf),
               o_brow = place(s.blk_row.data(), 4 * s.blk_row.size()), o_bcol = place(s.blk_col.data(), 4 * s.blk_col.size()),
               o_cdiag = place(s.col_ptr.data(), 4 * s.col_ptr.size()), o_prog = place(s.prog.data(), 4 * s.prog.size()),
               o_pptr = place(s.prog_ptr.data(), 4 * s.prog_ptr.size()), o_tprog = place(s.tree.words.data(), 4 * s.tree.words.size()),
               o_prodp = place(prod_ptr.data(), 4 * prod_ptr.size()), o_prod = place(prod.data(), 4 * prod.size());
  const size_t o_pose1 = place(nullptr, 56 * (size_t)n_poses), o_sysH = place(nullptr, 8 * P.sys_doubles), o_sys = place(nullptr, 8 * P.sys_doubles),
               o_xp = place(nullptr, 48 * (size_t)nf), o_scal = place(nullptr, 64), o_ctl = place(nullptr, sizeof(Control)),
               o_txchg = place(nullptr, 8 * (size_t)std::max(s.tree.xchg_doubles, 2)), o_stage = place(nullptr, 8 * 120 * (size_t)std::max(ne, 1));
  const size_t total = align_up(top);
  if (total > h->d_pg_bytes) {
    if (h->d_pg) cudaFree(h->d_pg);
    h->d_pg = nullptr; h->d_pg_bytes = 0;
    if (cudaMalloc((void **)&h->d_pg, total + total / 4) != cudaSuccess) { cudaGetLastError(); return fail(h, SSBA_ERR_ALLOC, "cudaMalloc failed"); }
    h->d_pg_bytes = total + total / 4;
  }
  char *d = h->d_pg;
  for (auto &it : items)
    if (it.src && it.bytes) CUDA_TRY(h, cudaMemcpyAsync(d + it.off, it.src, it.bytes, cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(h, cudaMemcpyAsync(d + o_pose1, d + o_pose0, 56 * (size_t)n_poses, cudaMemcpyDeviceToDevice, h->stream));
  P.pose[0] = (double *)(d + o_pose0); P.pose[1] = (double *)(d + o_pose1);
  P.pose_of_q = (const int32_t *)(d + o_poq); P.blk_row = (const int32_t *)(d + o_brow); P.blk_col = (const int32_t *)(d + o_bcol);
  P.col_diag = (const int32_t *)(d + o_cdiag); P.prog = (const int32_t *)(d + o_prog); P.prog_ptr = (const int32_t *)(d + o_pptr);
  P.sys = (double *)(d + o_sys); P.xp = (double *)(d + o_xp); P.scal = (double *)(d + o_scal); P.ctl = (Control *)(d + o_ctl);
  fill_tree_dev(s.tree, (const int32_t *)(d + o_tprog), (double *)(d + o_txchg), P.tree);
  const int32_t *d_ev0 = (const int32_t *)(d + o_ev0), *d_ev1 = (const int32_t *)(d + o_ev1), *d_eq0 = (const int32_t *)(d + o_eq0),
                *d_eq1 = (const int32_t *)(d + o_eq1), *d_eblk = (const int32_t *)(d + o_eblk);
  const double *d_minv = (const double *)(d + o_minv);
  double *d_sysH = (double *)(d + o_sysH);
  // ---- LM: the same device-side control as run_lm
  Control &c = *h->h_ctl;
  std::memset(&c, 0, sizeof(c));
  c.tau = h->opt.tau; c.good_lower = h->opt.good_step_lower_scale; c.good_upper = h->opt.good_step_upper_scale;
  c.user_lambda = h->opt.user_lambda_init; c.max_trials = h->opt.max_trials_after_failure;
  c.lambda = -1.0; c.ni = 2.0; c.cur = 0; c.need_linearize = 1; c.first_iteration = 1;
  c.max_iters = iters; c.last_result = SSBA_SOLVER_OK; c.world = 1; c.rank = 0;
  if (iters == 0) c.done = 1;
  CUDA_TRY(h, cudaMemcpyAsync(P.ctl, &c, sizeof(Control), cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  bool first = true;
  int guard = 0;
  const int max_slots = iters * (c.max_trials > 0 ? c.max_trials : 1) + 1;
  while (!c.done) {
    const int batch = std::max(1, iters - c.outer_iter);
    for (int i = 0; i < batch; ++i) {
      launch_pose_graph_slot(P, ne, d_ev0, d_ev1, d_eq0, d_eq1, d_eblk, d_minv, d_sysH, (double *)(d + o_stage),
                             (const int32_t *)(d + o_prodp), (const int32_t *)(d + o_prod), first, h->stream);
      h->prof.kernel_launches += first ? 6 : 5;
      first = false;
    }
    CUDA_TRY(h, cudaMemcpyAsync(&c, P.ctl, sizeof(Control), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    CUDA_TRY(h, cudaGetLastError());
    guard += batch;
    if (guard > max_slots) return fail(h, SSBA_ERR_STATE, "pose_graph_optimize: LM driver did not terminate");
  }
  // ---- results
  launch_pose_graph_chi(P, ne, d_ev0, d_ev1, d_minv, h->stream);
  CUDA_TRY(h, cudaMemcpyAsync(h->h_small, P.scal, 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaMemcpyAsync(poses_out, P.pose[c.cur], 56 * (size_t)n_poses, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  if (report) {
    report->iterations = c.outer_iter; report->last_result = c.last_result; report->n_records = c.n_records;
    report->cholesky_failures = c.cholesky_failures; report->chi2_initial = c.chi2_initial;
    report->chi2_robust = report->chi2_plain = h->h_small[0]; report->lambda = c.lambda;
    std::memcpy(report->iters, c.records, sizeof(ssba_iter_record) * c.n_records);
    report->seconds_total = secs(t0, Clock::now());
  }
  return SSBA_OK;
}

ssba_status ssba_set_profiling(ssba_handle *h, int32_t on) {
  if (!h) return SSBA_ERR_INVALID_ARG;
  if (h->stream) cudaStreamSynchronize(h->stream);
  h->opt.profile = on ? 1 : 0;
  h->P.pdl = pdl_enabled(h) ? 1 : 0;
  h->spans.clear();
  h->ev_used = 0;
  return SSBA_OK;
}

ssba_status ssba_profile_get(ssba_handle *h, ssba_profile *out) {
  if (!h || !out) return SSBA_ERR_INVALID_ARG;
  h->prof.cholesky_nnz = h->initialized ? 36LL * (h->s.n_blocks - h->s.n_fp) + 21LL * h->s.n_fp : 0;
  h->prof.hessian_pose_dimension = h->initialized ? 6 * h->s.n_fp : 0;
  h->prof.hessian_landmark_dimension = h->initialized ? 3 * h->s.n_fl_global : 0;
  h->prof.ms_symbolic_decomposition = h->ms_symbolic;
  h->prof.ms_structure_build = h->ms_structure_build;
  h->prof.ms_numeric_decomposition = h->prof.ms_reduced_solve;
  *out = h->prof;
  return SSBA_OK;
}

ssba_status ssba_profile_reset(ssba_handle *h) {
  if (!h) return SSBA_ERR_INVALID_ARG;
  std::memset(&h->prof, 0, sizeof(h->prof));
  return SSBA_OK;
}

ssba_status ssba_get_problem_info(ssba_handle *h, ssba_problem_info *out) {
  if (!h || !out) return SSBA_ERR_INVALID_ARG;
  if (!h->initialized) return fail(h, SSBA_ERR_STATE, "problem_info: not initialised");
  out->n_free_poses = h->s.n_fp; out->n_free_points = h->s.n_fl_global;
  out->n_active_edges = h->s.n_active_edges_global; out->n_pairs = h->s.n_pairs;
  out->n_schur_blocks = h->s.n_schur_blocks; out->n_factor_blocks = h->s.n_blocks;
  out->device_bytes = (int64_t)h->device_bytes;
  out->solve_cluster = h->s.solve_cluster;
  out->peer_exchange = h->use_p2p ? 1 : 0;
  out->solver_kind = h->s.tree.ok ? 1 : 0;
  out->solver_steps = h->s.tree.ok ? h->s.tree.chain_steps : h->s.n_levels;
  out->solver_top_cols = h->s.tree.ok ? h->s.tree.n_top_cols : 0;
  out->solver_smem_bytes = h->s.tree.ok ? (int32_t)h->s.tree.smem_bytes : 0;
  out->n_structure_builds = h->n_structure_builds; out->n_structure_reuses = h->n_structure_reuses;
  return SSBA_OK;
}

}  // extern "C"
