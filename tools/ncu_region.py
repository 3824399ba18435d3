"""Stall-reason breakdown of an address range of source lines (file, lo..hi) of one kernel in an ncu report.
usage: python tools/ncu_region.py <rep> <mangled kernel> <cubin tag> <file> <line_lo> <line_hi>"""
import collections, csv, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ncu_lines

rep, kern, tag, fn, lo, hi = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], int(sys.argv[5]), int(sys.argv[6])
lm = ncu_lines.line_map(tag, kern)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hint = os.environ.get("NCU_KERNEL", "tube")
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
for a, b in zip(starts[:-1], starts[1:]):
    if hint in rows[a][1]:
        rows = rows[a:b]
        break
h, data = rows[1], rows[2:]
ix = {n: i for i, n in enumerate(h)}
base = int(data[0][ix["Address"]], 16)
stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
agg = collections.Counter()
ins = smp = 0.0
ops = collections.Counter()
for r in data:
    off = int(r[ix["Address"]], 16) - base
    f, ln = lm.get(off, ("?", 0))
    if f != fn or not (lo <= ln <= hi):
        continue
    e = float(r[ix["Instructions Executed"]] or 0)
    ins += e
    smp += float(r[ix["# Samples"]] or 0)
    for s in stalls:
        agg[s] += float(r[ix[s]] or 0)
    op = [o for o in r[ix["Source"]].split() if not o.startswith("@")]
    if op:
        ops[op[0]] += e
print("instr %.4e samples %.0f" % (ins, smp))
tot = sum(agg.values()) or 1
print("stalls: " + ", ".join("%s %.1f%%" % (k[6:], 100 * v / tot) for k, v in agg.most_common(8)))
print("ops: " + ", ".join("%s %.1f%%" % (k, 100 * v / ins) for k, v in ops.most_common(16)))
