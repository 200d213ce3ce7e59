#!/usr/bin/env python
"""bench.py — LM iterations/s of ssvio's back-end local bundle adjustment on B200.

Metric (BASELINE.json): LM iterations/sec on the 100 key-frame / 20k landmark / 200k edge
local BA (KITTI-00 local-window shape, Huber delta 5.891), fp64, synthetic graph from
ssvio_b200.synth (seed 42).  One "step" = one optimize(10) = 10 outer LM iterations
(SparseOptimizer::optimize, thirdparty/g2o/g2o/core/sparse_optimizer.cpp:366-431) on that graph.

  value   whole-job LM it/s with the graph resident in HBM (reset of the estimates + 10 iterations)
  e2e     the same through the C ABI with HOST buffers, a NEW graph every step: set_* (H2D + structure build),
          optimize(10), get_poses/get_points (D2H); e2e.same_topology = the same call sequence when the
          indices are those of the resident structure (rounds 2..5 of backend.cpp:175-203: values only);
          e2e_shim (N = 1, when the compiled reference headers are available) = the real drop-in,
          g2o::SparseOptimizer + OptimizationAlgorithmLevenbergCuda (oracle/shim_harness.cpp)
  --impl reference   the reference's own g2o/CSparse CPU path (oracle/_ref, compiled from the
          reference tree; falls back to the C restatement when that .so is absent)

N > 1 (torchrun): the landmarks and their edges are sharded over the ranks, the reduced pose
system is all-reduced over NCCL each LM trial (strong scaling: the same graph).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "LM iterations/sec on 100KF/20k-pt/200k-edge BA"
UNIT = "LM it/s"
WORKLOAD = "cfg3"
WORKLOAD_DESC = "100 KF / 20k landmarks / 200k edges (KITTI-00 local-window shape), Huber 5.891, 10 LM iters"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ssba", choices=["ssba", "reference"])
    ap.add_argument("--workload", default=WORKLOAD)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-shim", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    return ap.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        """Started BEFORE the warm-up (nvidia-smi needs ~1 s to come up); mark() notes where the
        timed region begins so only samples taken under the timed load are kept."""
        self.n_before = 0
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-i", str(self.index), "-lms", "20"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def mark(self):
        try:
            self.f.flush()
            self.n_before = len(open(self.f.name).read().strip().splitlines())
        except Exception:
            self.n_before = 0

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        if len(rows) - self.n_before >= 3:
            rows = rows[self.n_before:]  # samples of the timed region only
            out["window"] = "timed region"
        else:
            out["window"] = "warm-up + timed region (timed region shorter than the sampling period)"
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for k, nme in enumerate(names):
                    if r[5 + k].strip().lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        if sm:
            sm.sort()
            out.update({"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                        "samples": len(sm)})
        return out


def iter_algorithmic_bytes(n_edges, n_points, n_poses, nnz_s):
    """bytes one LM iteration has to move (one linearise + one damped solve +
    one update + one chi2 pass), compact 28 B edge records."""
    return 3 * 28 * n_edges + 5 * 24 * n_points + 5 * 56 * n_poses + 2 * 288 * nnz_s + 2 * 8 * 6 * n_poses


# ------------------------------------------------------------------------------------------------
def run_reference(args):
    """The reference's own CPU implementation of the path, timed on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import bindings
    from ssvio_b200 import synth
    g = synth.make_config(args.workload)
    if bindings.RefOracle.available():
        orc, kind, jac = bindings.RefOracle(), "reference", "numeric"
        run = lambda it: orc.optimize(g, iters=it, jacobian=jac, trace=False, want_state=False)["report"]
    else:
        bindings.build(which=("port",))
        orc, kind, jac = bindings.PortOracle(), "port", "numeric"
        run = lambda it: orc.optimize(g, iters=it, jacobian=jac)["report"]
    # bounded sample: the step is the full optimize(iters) of the workload (same config as the CUDA arm); what is
    # bounded is the number of warm-up passes (a CPU run has nothing to warm) and, only if K full steps would
    # exceed ~4 minutes, the iterations per step (then said so in `sample`)
    rep0 = run(1)
    t_it = rep0.seconds_total  # optimize() only, no graph construction
    its = g.iters
    budget = 240.0
    if args.steps * its * t_it > budget:
        its = int(max(1, min(g.iters, budget / max(1e-9, args.steps * t_it))))
    warm = args.warmup if (args.steps + args.warmup) * its * t_it <= budget else 0
    for _ in range(warm):
        run(its)
    t_total, n_its = 0.0, 0
    for _ in range(args.steps):
        rep = run(its)
        t_total += rep.seconds_total  # wall time of optimize() only, as BASELINE.md defines it
        n_its += rep.iterations
    value = n_its / t_total
    sample = (f"{args.workload} full graph, optimize({its}) per step"
              + ("" if its == g.iters else f" ({its} of {g.iters} LM iterations: bounded)")
              + f", {warm} warm-up passes run, Jacobians as shipped (numeric), single thread "
                f"(G2O_USE_OPENMP OFF as the reference ships); host has {os.cpu_count()} cores")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(args.workload, g, its),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def run_ssba(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from ssvio_b200 import ba, build, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — libssba has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    multi = world > 1
    if multi:
        dist.init_process_group("nccl", device_id=dev)
    if rank == 0:
        build.build()
    if multi:
        dist.barrier()
    ba.load_library()

    nccl_id = None
    if multi:
        idt = torch.zeros(ba.SSBA_NCCL_ID_BYTES, dtype=torch.uint8, device=dev)
        if rank == 0:
            idt = torch.tensor(list(ba.nccl_unique_id()), dtype=torch.uint8, device=dev)
        dist.broadcast(idt, 0)
        nccl_id = bytes(idt.cpu().tolist())

    g = synth.make_config(args.workload)
    iters = g.iters
    stream = torch.cuda.current_stream().cuda_stream

    def barrier_sync():
        if multi:
            dist.barrier()
        torch.cuda.synchronize()

    # L2 flush between timed steps: the whole working set (~33 MB) would otherwise sit in the
    # 126 MB L2 from the previous step
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def flush_l2():
        flush.add_(1)

    # ---------------- resident: graph uploaded once, timed = reset + optimize(10)
    # several GPUs: pre-sharded input - every rank is handed (and uploads, and builds the structure of) only the
    # edges of the landmarks it owns; the partition itself (contiguous ranges of landmark rows with equal edge counts)
    # is the caller's bookkeeping, made once here
    opt = ba.BundleAdjuster(device_id=local_rank, stream=stream, rank=rank, world_size=world, nccl_id=nccl_id, presharded=multi)
    g_full = g
    if multi:
        g = g.shard(rank, world)
    opt.set_graph(g)
    opt.initialize_optimization()
    info = opt.problem_info()
    rep = opt.optimize(iters)  # also the numerics record of this run
    chi2_final, its_done = rep.chi2_robust, rep.iterations

    def step_resident():
        opt.reset_state()
        opt.optimize_nowait_report(iters)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # keep the GPUs busy ~1.5 s while nvidia-smi comes up; the steps are collective when N > 1,
    # so rank 0 decides and every rank follows
    t_s = time.perf_counter()
    while True:
        go = torch.tensor([1 if time.perf_counter() - t_s < 1.5 else 0], device=dev)
        if multi:
            dist.broadcast(go, 0)
        if int(go.item()) == 0:
            break
        flush_l2(); step_resident()
    for _ in range(max(args.warmup, 3)):
        flush_l2(); step_resident()
    barrier_sync()
    opt.profile_reset()
    if rank == 0:
        sampler.mark()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier_sync()
    t_wall0 = time.perf_counter()
    for a, b in ev:
        flush_l2()
        a.record(); step_resident(); b.record()
    barrier_sync()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if rank == 0 else None
    launches = opt.profile().kernel_launches
    ms_steps = [a.elapsed_time(b) for a, b in ev]
    total_ms = sum(ms_steps)
    if multi:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    value = iters * args.steps / (total_ms * 1e-3)

    # ---------------- per-phase device time (separate profiled pass: event pairs around phases,
    # recorded by the library on its launch stream; same handle, so the same NCCL communicator)
    opt.set_profiling(True)
    for _ in range(3):
        flush_l2(); step_resident()
    barrier_sync()
    opt.profile_reset()
    for _ in range(args.steps):
        flush_l2(); step_resident()
    barrier_sync()
    p = opt.profile()
    opt.set_profiling(False)
    phases = {"linearize": (p.ms_linearize, p.n_linearize), "schur": (p.ms_schur, p.n_schur),
              "reduced_solve": (p.ms_reduced_solve, p.n_reduced_solve),
              "update_chi2": (p.ms_update_chi2, p.n_update_chi2)}
    if multi:
        phases["allreduce"] = (p.ms_allreduce, p.n_allreduce)

    # ---------------- e2e: host buffers in, host buffers out, every step
    def pin(a):
        return torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    hg = synth.Graph(K=pin(g.K), ext=pin(g.ext), poses=pin(g.poses), pose_fixed=pin(g.pose_fixed),
                     points=pin(g.points), point_fixed=pin(g.point_fixed), pose_idx=pin(g.pose_idx),
                     point_idx=pin(g.point_idx), cam_idx=pin(g.cam_idx), uv=pin(g.uv),
                     huber_delta=g.huber_delta, iters=iters)
    eopt = ba.BundleAdjuster(device_id=local_rank, stream=stream) if not multi else opt
    e2e_chi = None
    e2e_out = None

    def step_e2e(fresh):
        nonlocal e2e_chi, e2e_out
        if fresh:
            eopt.drop_structure()            # a NEW window: the structure is built again (host) and uploaded
        eopt.set_graph(hg)                   # H2D of every input (+ structure build unless the topology is resident)
        r = eopt.optimize(iters)             # incl. final chi2 read-back (the step's result)
        e2e_chi = r.chi2_robust
        poses = eopt.poses()                 # D2H
        points = eopt.points()               # D2H (all-reduced gather of the shards when N > 1)
        e2e_out = (poses, points)

    def time_e2e(fresh):
        for _ in range(max(args.warmup, 3)):
            step_e2e(fresh)
        barrier_sync()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for a, b in evs:
            flush_l2()
            a.record(); step_e2e(fresh); b.record()
        barrier_sync()
        ms = sum(a.elapsed_time(b) for a, b in evs)
        if multi:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    h2d_total = hg.input_bytes()
    if multi:  # the ranks upload different amounts: the sum over the ranks
        t = torch.tensor([float(h2d_total)], dtype=torch.float64, device=dev)
        dist.all_reduce(t)
        h2d_total = int(t.item())
    e2e_ms = time_e2e(True)
    builds = eopt.problem_info().n_structure_builds
    same_ms = time_e2e(False)
    reuses = eopt.problem_info().n_structure_reuses
    e2e = {"value": iters * args.steps / (e2e_ms * 1e-3), "unit": UNIT,
           "h2d_bytes_per_step": h2d_total, "d2h_bytes_per_step": (hg.output_bytes() + 32) * world,
           "ms_per_step": e2e_ms / args.steps, "chi2_robust": e2e_chi,
           "what": "new window every step: set_* + structure build + upload + optimize(%d) + read-back" % iters,
           "structure_builds_counted": int(builds),
           "same_topology": {"value": iters * args.steps / (same_ms * 1e-3), "unit": UNIT, "ms_per_step": same_ms / args.steps,
                             "what": "the indices of the resident structure again (rounds 2..5 of backend.cpp:175-203): values only",
                             "structure_reuses_counted": int(reuses)}}
    # several GPUs: the parity facts of this very run (the driver's GPU test box has one GPU)
    parity_n = None
    if multi:
        import hashlib
        hsh = hashlib.sha256(e2e_out[0].tobytes() + e2e_out[1].tobytes()).digest()
        ht = torch.tensor(list(hsh), dtype=torch.uint8, device=dev)
        allh = [torch.zeros_like(ht) for _ in range(world)]
        dist.all_gather(allh, ht)
        parity_n = {"ranks_bitwise_equal": bool(all(bool((x == allh[0]).all()) for x in allh))}
    if not multi:
        eopt.close()

    # ---------------- the real drop-in (N = 1): g2o::SparseOptimizer + OptimizationAlgorithmLevenbergCuda
    # (include/ssba_g2o_shim.hpp) through the harness compiled against the reference's own headers
    e2e_shim = None
    if rank == 0 and not multi and not args.no_shim:
        try:
            from oracle import bindings
            if os.path.exists(bindings.SHIM_SO):
                shim = bindings.ShimHarness()
                e2e_shim = {}
                for key, every in (("write_back_once", False), ("write_back_every_iteration", True)):
                    shim.set_write_back_every_iteration(every)
                    for _ in range(3):
                        shim.optimize(g)
                    n = max(3, min(args.steps, 20))
                    secs, secs_init = [], []
                    for _ in range(n):
                        flush_l2(); torch.cuda.synchronize()
                        r = shim.optimize(g)["report"]
                        secs.append(r.seconds_total); secs_init.append(r.seconds_setup)
                    e2e_shim[key] = {"value": iters / (sum(secs) / n), "unit": UNIT, "ms_per_step": 1e3 * sum(secs) / n, "chi2_robust": r.chi2_robust,
                                     "g2o_initializeOptimization_ms": 1e3 * sum(secs_init) / n}
                e2e_shim["what"] = ("optimizer.optimize(%d) on a g2o::SparseOptimizer built like backend.cpp:81-168 with the shim as its algorithm: "
                                    "host wall clock around optimize() = flattening the active graph (init()), structure build, upload, the LM "
                                    "iterations, the write-back into the g2o vertices and edges - the same span the reference arm times; "
                                    "g2o's own initializeOptimization() (sparse_optimizer.cpp:201-272, identical in both arms) is given beside it" % iters)
        except Exception as ex:  # the harness needs the reference headers at build time; optional
            e2e_shim = {"unavailable": str(ex)[:200]}

    # ---------------- CPU baseline on the box's host cores (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and not multi and not args.no_cpu_baseline:
        from oracle import bindings
        if bindings.RefOracle.available():
            orc, kind = bindings.RefOracle(), "reference"
            r = orc.optimize(g, iters=iters, jacobian="numeric", trace=False, want_state=False)["report"]
            cpu_chi = orc.optimize(g, iters=iters, jacobian="numeric", trace=True, want_state=False)["report"].chi2_robust
        else:
            bindings.build(which=("port",))
            orc, kind = bindings.PortOracle(), "port"
            r = orc.optimize(g, iters=iters, jacobian="numeric")["report"]
            cpu_chi = r.chi2_robust
        cpu = {"value": r.iterations / r.seconds_total, "unit": UNIT, "cores": 1, "kind": kind,
               "sample": f"{args.workload} full graph, one optimize({iters}) ({r.seconds_total:.2f} s), numeric Jacobians as shipped, "
                         f"single thread (reference ships G2O_USE_OPENMP OFF); host has {os.cpu_count()} cores",
               "chi2_robust": cpu_chi, "chi2_rel_err_gpu_vs_cpu": abs(chi2_final - cpu_chi) / cpu_chi}

    secondary = None
    if rank == 0 and not multi and not args.no_secondary:
        secondary = run_secondary(local_rank, stream)

    if rank == 0:
        peak, peak_src = measured_peaks()
        nnz_s = info.n_schur_blocks
        gl = g          # this rank's shard (kernel-level algorithmic bytes: what THIS GPU's launches move)
        g = g_full      # the workload (config, whole-iteration bytes)
        b_iter = iter_algorithmic_bytes(g.n_edges, g.n_points, g.n_poses, nnz_s)
        roofline, roofline_all = None, []
        if phases:
            alg_all = {  # algorithmic bytes per launch (DESIGN.md "Kernels"), whole graph
                "linearize": 28 * gl.n_edges + 24 * g.n_points // world + 56 * g.n_poses + 144 * info.n_pairs + 72 * g.n_points // world,
                "schur": 144 * info.n_pairs + 72 * g.n_points // world + 288 * nnz_s + 48 * g.n_poses,
                "reduced_solve": 2 * 288 * nnz_s + 2 * 48 * g.n_poses,
                "update_chi2": 28 * gl.n_edges + 144 * info.n_pairs + 2 * 24 * g.n_points // world + 56 * g.n_poses,
            }
            ncu = ncu_record() if args.workload == WORKLOAD and not multi else {}
            for k in alg_all:  # one entry per kernel of the LM trial; durations measured live (CUDA events of this run)
                ms_k, n_k = phases[k]
                dur_k = ms_k / max(n_k, 1) * 1e-3
                rec = dict(ncu.get(k.replace("_chi2", ""), {}))
                if k == "schur" and "schur_reduce" in ncu:  # the phase is two kernels: k_schur and k_schur_reduce
                    rec["dram_bytes"] = rec.get("dram_bytes", 0) + ncu["schur_reduce"].get("dram_bytes", 0)
                roofline_all.append({"kernel": k, "bound": "hbm", "achieved": alg_all[k] / dur_k / 1e9, "peak": peak, "unit": "GB/s",
                                     "frac": alg_all[k] / dur_k / 1e9 / peak, "algorithmic_bytes_per_launch": alg_all[k],
                                     "avg_launch_ms": ms_k / max(n_k, 1), "share_of_step": ms_k / max(1e-12, sum(v[0] for v in phases.values())),
                                     "traffic": rec.get("dram_bytes"), "l2_pct": rec.get("l2_pct"), "fp64_pipe_pct": rec.get("fp64_pipe_pct"),
                                     "issue_active_pct": rec.get("issue_active_pct"), "ncu_source": ncu.get("source")})
            # dominant kernel = the phase with the largest share of the step
            name = max((k for k in phases if k != "allreduce"), key=lambda k: phases[k][0])
            dom = next(r for r in roofline_all if r["kernel"] == name)
            roofline = {"kernel": name, "bound": "hbm", "achieved": dom["achieved"], "peak": peak, "unit": "GB/s",
                        "frac": dom["frac"], "traffic": dom["traffic"], "peak_source": peak_src,
                        "algorithmic_bytes_per_launch": dom["algorithmic_bytes_per_launch"], "avg_launch_ms": dom["avg_launch_ms"],
                        "note": ("the reduced solve is bound by its chain of dependent 6x6 pivot blocks (latency), not by bytes: see DESIGN.md"
                                 if name == "reduced_solve" else
                                 "per-pair fp64 chains and L2 round trips at 12 warps per SM (latency), not bytes: see DESIGN.md"),
                        "phase_ms_per_step": {k: v[0] / args.steps for k, v in phases.items()}}
        golden = golden_chi2(args.workload)
        numerics = {"chi2_robust_final": chi2_final, "lm_iterations_done": its_done,
                    "chi2_reference_golden": golden,
                    "chi2_rel_err_vs_golden": (abs(chi2_final - golden) / golden) if golden else None}
        if parity_n:
            numerics.update(parity_n)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {**workload_config(args.workload, g, iters),
                       "l2": "flushed between timed steps (256 MiB write); working set %.1f MB < 126 MB L2" % (info.device_bytes / 1e6),
                       "parallelism": ("landmark-sharded x%d (pre-sharded input: every rank is handed its own landmarks' edges), %s of the reduced pose system" % (world, "NVLink peer-memory exchange" if info.peer_exchange else "NCCL all-reduce")) if multi else "single GPU",
                       "iter_algorithmic_bytes": b_iter,
                       "step_hbm_frac": (b_iter * iters / (total_ms / args.steps * 1e-3)) / 1e9 / peak,
                       "chi2_robust_final": chi2_final, "lm_iterations_done": its_done,
                       "wall_s_timed_region": t_wall},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "roofline_all": roofline_all, "cpu_baseline": cpu,
            "numerics": numerics, "e2e_shim": e2e_shim,
            "solver": {"kind": "subtree-per-CTA block LDL^T (k_tree_solve)" if info.solver_kind == 1 else "level-scheduled (k_reduced_solve)",
                       "cluster": info.solve_cluster, "chain_steps": info.solver_steps, "top_columns": info.solver_top_cols,
                       "smem_bytes": info.solver_smem_bytes},
            "secondary": secondary,
        }
        print(json.dumps(line), flush=True)
    opt.close()
    if multi:
        dist.barrier()
        dist.destroy_process_group()


def run_secondary(device_id, stream):
    """The widened rows of SURVEY 8f next to the reference's own CPU code, bounded to a few seconds: the batched
    pose-only LM of the front-end (frames/s, host buffers in and out) and the loop closer's pose-graph
    optimisation (ms per optimize(20) on a 400 key-frame graph)."""
    from ssvio_b200 import ba, synth
    out = {}
    try:
        from oracle import bindings
        ref = bindings.RefOracle() if bindings.RefOracle.available() else None
    except Exception:
        ref = None
    with ba.BundleAdjuster(device_id=device_id, stream=stream) as opt:
        nf = 4096
        b = synth.make_pose_only(nf, 150, seed=1)
        for _ in range(3):
            opt.pose_only_optimize(b)
        t = time.perf_counter()
        reps = 10
        for _ in range(reps):
            res = opt.pose_only_optimize(b)
        dt = (time.perf_counter() - t) / reps
        po = {"metric": "pose-only LM (FrontEnd::EstimateCurrentPose), frames/s", "value": nf / dt, "ms_per_batch": 1e3 * dt,
              "frames": nf, "features_per_frame": 150, "what": "host buffers in and out, 4 rounds x optimize(10), one launch"}
        if ref is not None:
            n_ref = 256
            sub = synth.PoseOnlyBatch(K=b.K, feat_ptr=b.feat_ptr[:n_ref + 1], poses=b.poses[:n_ref], xyz=b.xyz[:b.feat_ptr[n_ref]], uv=b.uv[:b.feat_ptr[n_ref]])
            t = time.perf_counter(); r = bindings.ref_pose_only(ref.lib, sub); dt_ref = time.perf_counter() - t
            po["cpu_reference"] = {"value": n_ref / dt_ref, "cores": 1, "sample": f"{n_ref} of the {nf} frames, g2o + LinearSolverDense as the reference"}
            po["inlier_decisions_equal_on_sample"] = bool((r[2] == res[2][:n_ref]).all())
        out["pose_only"] = po
        pg = synth.make_pose_graph(400, seed=9, n_loops=8)
        for _ in range(3):
            opt.pose_graph_optimize(pg)
        t = time.perf_counter()
        reps = 10
        for _ in range(reps):
            poses, rep = opt.pose_graph_optimize(pg)
        dt = (time.perf_counter() - t) / reps
        pgr = {"metric": "pose-graph optimisation (LoopClosing::PoseGraphOptimization), ms per optimize(20)", "value": 1e3 * dt,
               "key_frames": 400, "edges": int(len(pg.v0)), "chi2": rep.chi2_robust, "higher_is_better": False}
        if ref is not None:
            t = time.perf_counter(); rp, rr = bindings.ref_pose_graph(ref.lib, pg); dt_ref = time.perf_counter() - t
            pgr["cpu_reference"] = {"value": 1e3 * dt_ref, "cores": 1, "chi2": rr.chi2_robust, "sample": "the same graph, g2o + LinearSolverEigen as the reference"}
        out["pose_graph"] = pgr
    return out


def workload_config(name, g, iters):
    """The keys both arms share (the driver compares the arms' configs)."""
    return {"workload": workload_desc(name, g), "n_poses": g.n_poses, "n_points": g.n_points, "n_edges": g.n_edges,
            "lm_iters_per_step": iters, "seed": 42}


def workload_desc(name, g):
    if name == WORKLOAD:
        return WORKLOAD_DESC
    return f"{name}: {g.n_poses} KF / {g.n_points} landmarks / {g.n_edges} edges, Huber {g.huber_delta}, {g.iters} LM iters (not the headline workload)"


def ncu_record():
    """Per-kernel numbers of the committed `ncu --set full` capture of this workload (profiles/r2_ncu_kernels.json,
    written by scripts/ncu_summary.py --json from the .ncu-rep of the same bench command): DRAM bytes per launch,
    L2 / fp64-pipe / issue-active percentages.  These are profile constants, not measured in this run."""
    for name in ("r2_ncu_kernels.json",):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                return json.load(f)
        except (OSError, ValueError):
            pass
    return {}


def golden_chi2(workload):
    """Final robust chi2 of the compiled reference (as shipped: numeric Jacobians) on this config, from the committed
    fixtures (tests/golden/scalars.json, made by tests/golden/make_golden.py)."""
    try:
        with open(os.path.join(ROOT, "tests", "golden", "scalars.json")) as f:
            d = json.load(f)
        return float(d[workload]["numeric"]["chi2_robust"])
    except (OSError, ValueError, KeyError, TypeError):
        return None


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ssba(args)


if __name__ == "__main__":
    main()
