"""Attributes the per-instruction counts of an ncu report to CUDA source lines, using nvdisasm line info of the
in-tree library.  usage: python tools/ncu_lines.py <rep> <kernel mangled name> <cubin tag: kernels_f64|kernels_f32> [blocks]"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def line_map(tag, kernel):
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "gnuspeech_b200", "lib", "libtrm_cuda.so")], cwd=d,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    cub = [f for f in os.listdir(d) if f.startswith(tag + ".") and f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cub)], stdout=subprocess.PIPE, text=True).stdout
    m, cur, on = {}, ("?", 0), False
    for ln in txt.splitlines():
        if ln.startswith("//---") and ".text." in ln:
            on = (".text." + kernel + " ") in ln or ln.rstrip("- ").endswith(kernel)
            continue
        if not on:
            continue
        mm = re.match(r'\s*//## File "(.*)", line (\d+)', ln)
        if mm:
            cur = (os.path.basename(mm.group(1)), int(mm.group(2)))
            continue
        mm = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
        if mm:
            m[int(mm.group(1), 16)] = cur
    return m


def main():
    rep, kernel, tag = sys.argv[1], sys.argv[2], sys.argv[3]
    B = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
    lm = line_map(tag, kernel)
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    # a report may hold several kernels: keep the section whose "Kernel Name" row matches the demangled hint
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
        off = int(r[ix["Address"]], 16) - base
        key = lm.get(off, ("?", 0))
        a = agg.setdefault(key, [0.0, 0.0])
        a[0] += float(r[ix["Instructions Executed"]] or 0)
        a[1] += float(r[ix["# Samples"]] or 0)
    ti = sum(a[0] for a in agg.values())
    ts = sum(a[1] for a in agg.values())
    src = {}
    for (fn, ln) in agg:
        if fn not in src:
            for p in (os.path.join(ROOT, "gnuspeech_b200", "csrc", fn),):
                if os.path.exists(p):
                    src[fn] = open(p).read().splitlines()
    print("total instr/block %.0f" % (ti / B))
    for (fn, ln), a in sorted(agg.items(), key=lambda kv: (kv[0][0] != "tube_wide.cuh", kv[0][0], kv[0][1])):
        if a[0] / ti < 0.002:
            continue
        text = src[fn][ln - 1].strip()[:90] if fn in src and 0 < ln <= len(src[fn]) else ""
        print("%-22s %4d  instr/blk %6.0f %5.1f%%  smp %5.1f%%  %s" % (fn[:22], ln, a[0] / B, 100 * a[0] / ti, 100 * a[1] / ts, text))


if __name__ == "__main__":
    main()
