"""Summarises an .ncu-rep (read here, without a GPU): headline metrics, stall reasons and the instruction mix
by loop multiplicity for the SASS of one kernel.  usage: python tools/ncu_summary.py <rep> [blocks_per_launch]"""
import collections
import csv
import subprocess
import sys


def page(rep, name, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"] + list(extra), stdout=subprocess.PIPE, text=True).stdout
    return list(csv.reader(out.splitlines()))


def main():
    rep = sys.argv[1]
    det = page(rep, "details")
    h = det[0]
    keep = ("Duration", "Registers Per Thread", "Theoretical Occupancy", "Achieved Occupancy", "Executed Ipc", "Issue Slots Busy",
            "Compute (SM) Throughput", "DRAM Throughput", "Dynamic Shared Memory Per Block", "Block Size", "Grid Size",
            "Avg. Active Threads", "Warp Cycles Per Issued", "Eligible Warps", "No Eligible", "L1/TEX Hit")
    print("== %s" % rep)
    for r in det[1:]:
        d = dict(zip(h, r))
        if any(k in d.get("Metric Name", "") for k in keep):
            print("%-45s %-16s %s" % (d["Metric Name"], d["Metric Unit"], d["Metric Value"]))
    raw = page(rep, "raw")
    for n, u, v in zip(raw[0], raw[1], raw[2]):
        if n in ("dram__bytes_read.sum", "dram__bytes_write.sum", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
                 "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
                 "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
                 "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "gpu__time_duration.sum",
                 "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"):
            print("%-70s %-10s %s" % (n, u, v))
    rows = page(rep, "source")
    import os
    hint = os.environ.get("NCU_KERNEL", "tube_wide")
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
    for a, b in zip(starts[:-1], starts[1:]):
        if hint in rows[a][1]:
            rows = rows[a:b]
            break
    h = rows[1]
    data = rows[2:]
    ix = {n: i for i, n in enumerate(h)}

    def f(r, n):
        try:
            return float(r[ix[n]])
        except Exception:
            return 0.0
    tot_s = sum(f(r, "# Samples") for r in data)
    tot_i = sum(f(r, "Instructions Executed") for r in data)
    print("warp instructions executed %.4e   stall samples %.0f" % (tot_i, tot_s))
    stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
    agg = sorted(((sum(f(r, n) for r in data), n) for n in stalls), reverse=True)
    print("stalls: " + ", ".join("%s %.1f%%" % (n[6:], 100 * v / tot_s) for v, n in agg[:8]))
    if len(sys.argv) > 2:
        B = float(sys.argv[2])
        cls = collections.OrderedDict()
        for r in data:
            e = f(r, "Instructions Executed")
            k = e / B
            key = "<0.5" if k < 0.5 else ("1 (per block)" if k < 1.5 else ("2-7" if k < 7.5 else ("8-15" if k < 15.5 else ("16" if k < 16.5 else ">16"))))
            c = cls.setdefault(key, [0.0, 0.0, 0])
            c[0] += e
            c[1] += f(r, "# Samples")
            c[2] += 1
        for k, c in cls.items():
            print("  mult %-14s sass rows %4d  instr/block %7.0f (%4.1f%%)  samples %4.1f%%" % (k, c[2], c[0] / B, 100 * c[0] / tot_i, 100 * c[1] / tot_s))
        ops = collections.Counter()
        for r in data:
            op = r[ix["Source"]].split()
            op = [o for o in op if not o.startswith("@")]
            if op:
                ops[op[0].split(".")[0]] += f(r, "Instructions Executed")
        print("  top opcodes/block: " + ", ".join("%s %.0f" % (o, v / B) for o, v in ops.most_common(24)))


if __name__ == "__main__":
    main()
