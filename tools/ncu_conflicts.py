"""Shared-memory wavefronts (ideal vs. excessive) per CUDA source line of one kernel of an ncu report.
usage: NCU_KERNEL=<name hint> python tools/ncu_conflicts.py <rep> <kernel mangled name> <cubin tag>"""
import collections
import csv
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ncu_lines import line_map


def main():
    rep, kernel, tag = sys.argv[1:4]
    lm = line_map(tag, kernel)
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hint = os.environ.get("NCU_KERNEL", "tube_wide")
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
    for a, b in zip(starts[:-1], starts[1:]):
        if hint in rows[a][1]:
            rows = rows[a:b]
            break
    h, data = rows[1], rows[2:]
    ix = {n: i for i, n in enumerate(h)}
    base = int(data[0][ix["Address"]], 16)
    agg = collections.OrderedDict()
    for r in data:
        w = float(r[ix["L1 Wavefronts Shared"]] or 0)
        if w == 0:
            continue
        key = lm.get(int(r[ix["Address"]], 16) - base, ("?", 0))
        a = agg.setdefault(key, [0.0, 0.0, 0.0, ""])
        a[0] += w
        a[1] += float(r[ix["L1 Wavefronts Shared Excessive"]] or 0)
        a[2] += float(r[ix["Instructions Executed"]] or 0)
        a[3] = r[ix["Source"]].split()[0]
    tot = sum(a[0] for a in agg.values()); exc = sum(a[1] for a in agg.values())
    print("shared wavefronts %.3e, excessive %.3e (%.1f%%)" % (tot, exc, 100 * exc / max(tot, 1)))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
        print("%-22s %5d  wavefronts %.3e  excessive %.3e  inst %.3e  (%s)" % (k[0], k[1], a[0], a[1], a[2], a[3]))


if __name__ == "__main__":
    main()
