"""Hottest source lines of one kernel in an ncu report by stall samples, with each line's dominant stall reasons.
usage: python tools/ncu_hot.py <rep> <mangled kernel> <cubin tag> [top]   (NCU_KERNEL = demangled-name hint)"""
import collections, csv, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ncu_lines

rep, kern, tag = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
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
agg = collections.defaultdict(lambda: [0.0, 0.0, collections.Counter()])
for r in data:
    off = int(r[ix["Address"]], 16) - base
    key = lm.get(off, ("?", 0))
    a = agg[key]
    a[0] += float(r[ix["Instructions Executed"]] or 0)
    a[1] += float(r[ix["# Samples"]] or 0)
    for s in stalls:
        a[2][s[6:]] += float(r[ix[s]] or 0)
ts = sum(a[1] for a in agg.values()) or 1
ti = sum(a[0] for a in agg.values()) or 1
src = {}
for (fn, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    if fn not in src:
        p = os.path.join(ncu_lines.ROOT, "gnuspeech_b200", "csrc", fn)
        src[fn] = open(p).read().splitlines() if os.path.exists(p) else []
    text = src[fn][ln - 1].strip()[:70] if 0 < ln <= len(src[fn]) else ""
    st = ", ".join("%s %.0f%%" % (k, 100 * v / max(sum(a[2].values()), 1)) for k, v in a[2].most_common(3))
    print("%-16s %4d smp %5.1f%% instr %5.1f%%  [%s]  %s" % (fn[:16], ln, 100 * a[1] / ts, 100 * a[0] / ti, st, text))
