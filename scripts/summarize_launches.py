"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel:
   python scripts/summarize_launches.py gpurun_out/launches.csv > profiles/<name>.md"""
import collections
import csv
import re
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = collections.Counter(), collections.Counter()
    for r in data:
        if len(r) <= vi:
            continue
        name = re.sub(r"\(.*", "", r[ki]).split("::")[-1]
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else v * 1e3 if r[ui] == "ms" else v
        tot[name] += v
        cnt[name] += 1
    T = sum(tot.values())
    print(f"source: {path} (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised launches: compare shares)\n")
    print("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|")
    for k, v in tot.most_common():
        print(f"| {k} | {cnt[k]} | {v:.1f} | {v / cnt[k]:.2f} | {100 * v / T:.1f}% |")


if __name__ == "__main__":
    main(sys.argv[1])
