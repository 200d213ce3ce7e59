"""Per-kernel instruction evidence from the in-tree library (cuobjdump -sass ssvio_b200/lib/libssba.so):
which kernels use the bulk-copy / TMA unit (UBLKCP), mbarriers (SYNCS), cluster barriers (UCGABAR), distributed
shared memory stores (ST.*CLUSTER / STS remote), asynchronous copies (LDGSTS), programmatic dependent launch
(ACQBULK / PREEXIT = griddepcontrol.wait / launch_dependents), fp64 FMAs (DFMA), fp64 tensor-core MMAs (DMMA)
and fp64 reductions (RED/REDG/ATOMG .F64: none are left on the LM path with the deterministic accumulation).

    python scripts/sass_summary.py > profiles/r2_sass_summary.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ssvio_b200", "lib", "libssba.so")
PAT = collections.OrderedDict([
    ("UBLKCP (cp.async.bulk, TMA unit)", r"\bUBLKCP"), ("SYNCS (mbarrier)", r"\bSYNCS"), ("UCGABAR (cluster barrier)", r"\bUCGABAR"),
    ("LDGSTS (cp.async)", r"\bLDGSTS"), ("ACQBULK (griddepcontrol.wait)", r"\bACQBULK"), ("PREEXIT (launch_dependents)", r"\bPREEXIT"),
    ("DFMA", r"\bDFMA"), ("DMMA", r"\bDMMA"), ("RED/ATOM .F64", r"\b(RED|REDG|ATOM|ATOMG)\.[A-Z.]*F64"), ("MUFU.RCP64H / RSQ64H", r"MUFU\.(RCP64H|RSQ64H)"),
    ("LDS", r"\bLDS\b|\bLDS\."), ("instructions", r"^\s+/\*[0-9a-f]{4,}\*/"),
])


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    arch = re.search(r"arch = (sm_\w+)", sass)
    cur, counts = None, collections.OrderedDict()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(anonymous namespace\)::", "", name)
            name = re.sub(r"^void ", "", name).split("(")[0].replace("ssba::", "")
            cur = name
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        for key, pat in PAT.items():
            if re.search(pat, line):
                counts[cur][key] += 1
    keys = list(PAT.keys())
    print(f"# SASS evidence per kernel ({os.path.relpath(LIB, ROOT)}, {arch.group(1) if arch else 'sm_100a'}; `python scripts/sass_summary.py`)\n")
    print("| kernel | " + " | ".join(keys) + " |")
    print("|---|" + "---|" * len(keys))
    for k, c in counts.items():
        print(f"| `{k}` | " + " | ".join(str(c.get(x, 0)) for x in keys) + " |")
    tot = collections.Counter()
    for c in counts.values():
        tot.update(c)
    print("| **all** | " + " | ".join(str(tot.get(x, 0)) for x in keys) + " |")


if __name__ == "__main__":
    main()
