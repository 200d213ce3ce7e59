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


if __name__ == "__main__":
    main(sys.argv[1])
