import sys, re, collections, statistics
d = collections.defaultdict(list)
for line in sys.stdin:
    m = re.match(r"\[ssba\]\s+(.*?)\s+([0-9.]+) ms$", line.strip())
    if m: d[m.group(1)].append(float(m.group(2)))
    m2 = re.match(r"\[ssba\] initialize: (.*)$", line.strip())
    if m2:
        for part in m2.group(1).split(", "):
            mm = re.match(r"(.*?) ([0-9.]+) ms", part)
            if mm: d["init:" + mm.group(1)].append(float(mm.group(2)))
for k, v in d.items():
    print(f"{k:34s} median {statistics.median(v[3:] or v):.3f} ms  (n={len(v)})")
