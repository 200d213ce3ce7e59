"""Key metrics per kernel from an ncu --set full report:
   python scripts/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.md"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"),
    ("launch__occupancy_limit_registers", "occ_lim_regs"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved_occ_%"),
    ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
    ("lts__t_bytes.sum", "l2_bytes"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_%"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1_%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64_pipe_%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_%"),
    ("smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "stall_long_sb"),
    ("smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio", "stall_short_sb"),
    ("smsp__average_warp_latency_issue_stalled_barrier.ratio", "stall_barrier"),
    ("smsp__average_warp_latency_issue_stalled_wait.ratio", "stall_wait"),
    ("smsp__average_warp_latency_issue_stalled_lg_throttle.ratio", "stall_lg_throttle"),
    ("smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio", "stall_math_throttle"),
    ("smsp__average_warp_latency_issue_stalled_membar.ratio", "stall_membar"),
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    cols = [(hdr.index(m), n, m) for m, n in WANT if m in hdr]
    print(f"source: {path} (ncu --set full --clock-control none)\n")
    print("| kernel | " + " | ".join(n for _, n, _ in cols) + " |")
    print("|---|" + "---|" * len(cols))
    for r in data:
        name = r[ki].split("(")[0].split("::")[-1]
        vals = []
        for i, n, m in cols:
            v = r[i]
            try:
                v = f"{float(v.replace(',', '')):.4g} {units[i]}".strip()
            except ValueError:
                pass
            vals.append(v)
        print(f"| {name} | " + " | ".join(vals) + " |")


def to_json(path):
    """profiles/r2_ncu_kernels.json: what bench.py's roofline_all quotes per kernel (averages over the captured launches)."""
    import json
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, data = rows[0], rows[2:]
    ki = hdr.index("Kernel Name")
    col = {m: hdr.index(m) for m, _ in WANT if m in hdr}
    names = {"k_linearize": "linearize", "k_schur": "schur", "k_schur_reduce": "schur_reduce", "k_tree_solve": "reduced_solve", "k_reduced_solve": "reduced_solve", "k_update": "update"}
    acc = {}
    for r in data:
        kn = r[ki].split("(")[0].split("::")[-1].split("<")[0]
        if kn not in names:
            continue
        f = lambda m: float(r[col[m]].replace(",", "")) if m in col and r[col[m]] not in ("", "n/a") else None
        unit = rows[1]
        def bytes_of(m):
            v = f(m)
            if v is None:
                return 0.0
            u = unit[col[m]].lower()
            return v * (1e9 if u.startswith("g") else 1e6 if u.startswith("m") else 1e3 if u.startswith("k") else 1.0)
        a = acc.setdefault(names[kn], {"n": 0, "dram_bytes": 0.0, "l2_pct": 0.0, "fp64_pipe_pct": 0.0, "issue_active_pct": 0.0, "duration_us": 0.0})
        a["n"] += 1
        a["dram_bytes"] += bytes_of("dram__bytes_read.sum") + bytes_of("dram__bytes_write.sum")
        a["l2_pct"] += f("lts__throughput.avg.pct_of_peak_sustained_elapsed") or 0.0
        a["fp64_pipe_pct"] += f("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active") or 0.0
        a["issue_active_pct"] += f("smsp__issue_active.avg.pct_of_peak_sustained_active") or 0.0
    res = {"source": f"{path} (ncu --set full --clock-control none; cold caches, serialised launches)"}
    for k, a in acc.items():
        n = max(a.pop("n"), 1)
        res[k] = {m: (round(v / n) if m == "dram_bytes" else round(v / n, 3)) for m, v in a.items() if m != "duration_us"}
        res[k]["launches_averaged"] = n
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--json":
        to_json(sys.argv[2])
    else:
        main(sys.argv[1])
